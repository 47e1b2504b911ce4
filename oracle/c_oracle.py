"""ORACLE -- TEST INFRASTRUCTURE ONLY.  ctypes binding of ``oracle/tt_oracle.c``.

The C file restates the reference path *including* the numpy / scipy algorithms it delegates to
(numpy.gradient, RegularGridInterpolator, solve_ivp RK45); this module only loads it.  Because it needs
no Python in its inner loops it integrates one ray per ``solve_ivp`` problem (``batch=1``: every ray gets
its own adaptive step sequence) at tight tolerances in milliseconds, which is what lets the GPU parity
tests run at thousands of rays on 129^3 .. 257^3 cubes.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "tt_oracle.c")
LIB = os.path.join(HERE, "_build", "libtt_oracle.so")
C_LIGHT = 299792458.0

_lib = None


class _Field(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
                ("x", C.c_void_p), ("y", C.c_void_p), ("z", C.c_void_p),
                ("gx", C.c_void_p), ("gy", C.c_void_p), ("gz", C.c_void_p)]


def build(force=False):
    """gcc the restatement (oracle/Makefile); a no-op when the library is newer than its source."""
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        r = subprocess.run(["make", "-s", "-C", HERE] + (["-B"] if force else []), capture_output=True, text=True)
        if r.returncode != 0 or not os.path.exists(LIB):
            raise RuntimeError("building oracle/tt_oracle.c failed:\n" + r.stdout + r.stderr)
    return LIB


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        dp, fp = C.c_void_p, C.POINTER(_Field)
        lib.tto_calc_dndr.argtypes = [dp, C.c_int, C.c_int, C.c_int, dp, dp, dp, C.c_double, C.c_double,
                                      dp, dp, dp, dp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        lib.tto_calc_dndr.restype = C.c_int
        lib.tto_dndr.argtypes = [fp, dp, C.c_long, dp]
        lib.tto_dndr.restype = None
        lib.tto_solve_ivp_rk45.argtypes = [fp, dp, C.c_long, C.c_double, C.c_double, C.c_double, dp,
                                           C.POINTER(C.c_long), C.POINTER(C.c_long), dp, C.c_long]
        lib.tto_solve_ivp_rk45.restype = C.c_long
        lib.tto_solve.argtypes = [fp, dp, C.c_long, C.c_long, C.c_double, C.c_double, C.c_double, dp, C.c_int,
                                  C.POINTER(C.c_long)]
        lib.tto_solve.restype = C.c_longlong
        lib.tto_solve_ms.argtypes = [fp, dp, C.c_long, C.c_long, C.c_double, C.c_double, C.c_double, C.c_double, dp, C.c_int,
                                     C.POINTER(C.c_long)]
        lib.tto_solve_ms.restype = C.c_longlong
        lib.tto_ray_at_exit.argtypes = [dp, C.c_long, C.c_double, C.c_int, dp]
        lib.tto_ray_at_exit.restype = None
        lib.tto_max_threads.restype = C.c_int
        _lib = lib
    return _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def calc_dndr(ne, x, y, z, lwl=1053e-9, ne_max=1):
    """particle_tracker.py:227-237 -> dict(omega, nc, ne_nc, dndx, dndy, dndz)."""
    lib = load()
    ne, x, y, z = _f64(ne), _f64(x), _f64(y), _f64(z)
    assert ne.shape == (x.size, y.size, z.size)
    out = [np.empty(ne.shape) for _ in range(4)]
    om, nc = C.c_double(), C.c_double()
    rc = lib.tto_calc_dndr(_p(ne), x.size, y.size, z.size, _p(x), _p(y), _p(z), float(lwl), float(ne_max),
                           _p(out[0]), _p(out[1]), _p(out[2]), _p(out[3]), C.byref(om), C.byref(nc))
    if rc:
        raise ValueError("tto_calc_dndr: every axis needs at least 2 points" if rc == 1 else "out of memory")
    return dict(omega=om.value, nc=nc.value, ne_nc=out[0], dndx=out[1], dndy=out[2], dndz=out[3])


class GradientField:
    """The three interpolators of particle_tracker.py:239-241 and ``dndr`` (:243-256)."""

    def __init__(self, x, y, z, dndx, dndy, dndz):
        self.x, self.y, self.z = _f64(x), _f64(y), _f64(z)
        self.g = [_f64(dndx), _f64(dndy), _f64(dndz)]
        for g in self.g:
            assert g.shape == (self.x.size, self.y.size, self.z.size)
        self.c = _Field(self.x.size, self.y.size, self.z.size, _p(self.x), _p(self.y), _p(self.z),
                        _p(self.g[0]), _p(self.g[1]), _p(self.g[2]))

    def dndr(self, pos):
        pos = _f64(pos)
        out = np.empty_like(pos)
        load().tto_dndr(C.byref(self.c), _p(pos), pos.shape[1], _p(out))
        return out


def make_field(ne, x, y, z, lwl=1053e-9, ne_max=1):
    d = calc_dndr(ne, x, y, z, lwl, ne_max)
    return GradientField(x, y, z, d["dndx"], d["dndy"], d["dndz"])


def ray_at_exit(sf, extent, probing_direction="z"):
    sf = _f64(sf)
    rf = np.empty((4, sf.shape[1]))
    load().tto_ray_at_exit(_p(sf), sf.shape[1], float(extent), "xyz".index(probing_direction), _p(rf))
    return rf


def solve(field, s0, extent, probing_direction="z", rtol=1e-3, atol=1e-6, batch=None, threads=0, strict=True,
          max_step=np.inf):
    """Same contract as ``oracle.ref_numpy.solve``: (rf, sf, ray_rhs_evals).

    ``batch=None`` is the reference's ``ElectronCube.solve`` (one adaptive step sequence for the whole
    bundle, particle_tracker.py:317-330); ``batch=k`` integrates bundles of k rays independently
    (``threads`` POSIX threads, 0 = all cores), ``batch=1`` gives every ray its own step control.
    A bundle whose step size underflows (solve_ivp's status -1) raises, or with ``strict=False`` comes back NaN.
    ``max_step`` (seconds) is solve_ivp's option of that name; the reference leaves it at infinity.  See tt_oracle.c:
    the sharp reference of the parity tests uses one cell's transit time (``cell_transit_time``).
    """
    lib = load()
    s0 = _f64(s0)
    n = s0.shape[1]
    sf = np.empty_like(s0)
    T = np.sqrt(8.0) * float(extent) / C_LIGHT
    failed = C.c_long()
    evals = lib.tto_solve_ms(C.byref(field.c), _p(s0), n, int(batch or n), T, float(rtol), float(atol), float(max_step),
                             _p(sf), int(threads), C.byref(failed))
    if evals < 0:
        raise MemoryError("tto_solve: out of memory")
    if failed.value and strict:
        raise RuntimeError(f"tto_solve: step size underflow in {failed.value} bundle(s)")
    return ray_at_exit(sf, extent, probing_direction), sf, int(evals)


def cell_transit_time(x, y, z):
    """time light needs across the smallest cell: the ``max_step`` of the sharp reference"""
    return min(float(np.diff(np.asarray(a, dtype=np.float64)).min()) for a in (x, y, z)) / C_LIGHT


def solve_one_bundle(field, s0, extent, rtol=1e-3, atol=1e-6, t_hist=None):
    """One solve_ivp problem with diagnostics: (sf, nfev, accepted steps, rejected steps); ``t_hist`` (a float64
    array) receives the end times of the accepted steps."""
    lib = load()
    s0 = _f64(s0)
    sf = np.empty_like(s0)
    ns, nr = C.c_long(), C.c_long()
    T = np.sqrt(8.0) * float(extent) / C_LIGHT
    nfev = lib.tto_solve_ivp_rk45(C.byref(field.c), _p(s0), s0.shape[1], T, float(rtol), float(atol), _p(sf),
                                  C.byref(ns), C.byref(nr), _p(t_hist) if t_hist is not None else None,
                                  t_hist.size if t_hist is not None else 0)
    if nfev < 0:
        raise RuntimeError(f"tto_solve_ivp_rk45 failed ({nfev})")
    return sf, int(nfev), ns.value, nr.value


def max_threads():
    return int(load().tto_max_threads())
