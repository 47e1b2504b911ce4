"""Drop-in mirror of the reference's ``particle_tracker`` module (ElectronCube, dsdt) whose hot
path runs in hand-written sm_100a CUDA behind the C ABI of ``include/tt_b200.h``.

Reference: particle_tracking/particle_tracker.py (class ElectronCube :121-395, dsdt :398-419).
Same names, argument meaning and results; what changed underneath:

* ``calc_dndr``   -> one stencil kernel writing an interleaved (g_u, g_v, g_w, ne/nc) grid
* ``solve``       -> one fused fixed-step RK4 kernel (plane marching, Morton-ordered rays)
                     instead of scipy ``solve_ivp`` over three ``RegularGridInterpolator`` objects
* large arrays (``s0``, ``sf``, ``rf``) may live in HBM as :class:`DeviceArray` (numpy-compatible).

Extra, optional keyword arguments (reference-compatible defaults): ``dtype`` ("float32" |
"float64": arithmetic and grid element type), ``steps_per_cell`` (RK4 sub-planes per grid cell),
``sort_rays`` (Morton ordering).  There is no CPU fallback: without a CUDA device ``calc_dndr`` /
``solve`` raise.
"""
from __future__ import annotations

import ctypes as C
from datetime import datetime
from time import time

import numpy as np

from . import _lib
from ._lib import DeviceArray, TTError   # noqa: F401  (TTError re-exported)

c = 299792458.0                 # scipy.constants.c, particle_tracker.py:119
_NC_COEFF = 3.14207787e-4       # particle_tracker.py:228
# Verdet-type constant e^3 / (8 pi^2 eps0 me^2 c^3) (CODATA 2018): rotation = VERDET * lambda^2 * int ne B_par ds
_E, _EPS0, _ME = 1.602176634e-19, 8.8541878128e-12, 9.1093837015e-31
VERDET = _E**3 / (8 * np.pi**2 * _EPS0 * _ME**2 * c**3)        # = 2.6312e-13 rad / (T m^2) per m^2 of wavelength^2
_AXIS = {"x": 0, "y": 1, "z": 2}
UNIFORM_RTOL = 1e-9


def _axis_spacing(a, name):
    """(origin, mean spacing, uniform?) of a coordinate axis.  Axes whose node spacing deviates from
    the mean by more than UNIFORM_RTOL of the axis length are rectilinear: they take the
    ``tt_*_axes`` kernels (node coordinates on the device) instead of origin + spacing."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim != 1 or a.size < 2:
        raise ValueError(f"axis {name} must be a 1-D array with at least 2 points")
    h = (a[-1] - a[0]) / (a.size - 1)
    d = np.diff(a)
    if not h > 0 or not np.all(d > 0):
        raise ValueError(f"axis {name} must be strictly ascending")
    uniform = bool(np.max(np.abs(d - h)) <= UNIFORM_RTOL * abs(a[-1] - a[0]))
    return float(a[0]), float(h), uniform


def _uniform_spacing(a, name):
    o, h, uniform = _axis_spacing(a, name)
    if not uniform:
        raise NotImplementedError(f"axis {name} is not uniformly spaced; this entry point needs uniform axes")
    return o, h


class _GradInterp:
    """Stand-in for the reference's ``dnd?_interp`` RegularGridInterpolator objects
    (particle_tracker.py:239-241): callable on (N, 3) points, zero outside the cube."""

    def __init__(self, cube, comp):
        self._cube, self._comp = cube, comp

    def __call__(self, xi):
        xi = np.asarray(xi, dtype=np.float64)
        return self._cube.dndr(np.ascontiguousarray(xi.reshape(-1, 3).T))[self._comp].reshape(xi.shape[:-1])


class ElectronCube:
    """A class to hold and generate electron density cubes (particle_tracker.py:121-145)."""

    def __init__(self, x, y, z, *args, probing_direction=None, B_on=False, inv_brems=False, phaseshift=False,
                 dtype="float32", steps_per_cell=1, sort_rays=True, keep_sf=True, verbose=True, face_grid="auto"):
        """x, y, z: 1-D coordinate arrays (m); probing_direction 'x' | 'y' | 'z' (4th positional argument or
        keyword, default 'z'; :125-145).

        The reference's example scripts call an older signature ``ElectronCube(x, y, z, extent, B_on=,
        inv_brems=, phaseshift=, probing_direction=)`` (example_kitchensink.py:72): a number in the 4th
        position is accepted as that ``extent`` (it is implied by the axes: ``extent = axis.max()``)."""
        for a in args:
            if isinstance(a, str):
                if probing_direction is not None:
                    raise TypeError("probing_direction given twice")
                probing_direction = a
            elif not isinstance(a, (int, float, np.integer, np.floating)):
                raise TypeError("ElectronCube(x, y, z[, probing_direction | extent], ...)")
        if len(args) > 1:
            raise TypeError("ElectronCube takes at most 4 positional arguments")
        if probing_direction is None:
            probing_direction = "z"
        self.z, self.y, self.x = z, y, x
        self.extent_x = x.max()
        self.extent_y = y.max()
        self.extent_z = z.max()
        self.probing_direction = probing_direction
        # magnetised / absorbing extension (parity unpinned: call sites only in the reference)
        self.B_on, self.inv_brems, self.phaseshift = bool(B_on), bool(inv_brems), bool(phaseshift)
        self._B = self._Te = None
        self._Z = 1.0
        self._aux = None
        self.coulomb_log = None
        self.dtype = "float64" if _lib.dtype_code(dtype) == _lib.TT_F64 else "float32"
        self.steps_per_cell = int(steps_per_cell)
        self.sort_rays = bool(sort_rays)
        self.keep_sf = bool(keep_sf)
        self.verbose = bool(verbose)
        self._ne = None          # host array or device tensor as supplied
        self._grid = None        # device tensor [nw, nv, nu, 4]
        # face-coefficient grid of the production FP32 kernel (tt_build_face_grid / tt_trace_faces: 48 B per cell face,
        # built lazily by solve() from the gradient grid).  "auto": used when it applies (float32, uniform axes,
        # 1 step per cell, trajectory only) and fits next to everything else; True: required; False: never
        self.face_grid = face_grid
        self._faces = None
        self._faces_valid = False
        self._faces_aux = None
        self._faces_aux_key = None
        self._nodes = None       # device node coordinates (x, y, z) when the axes are not uniformly spaced
        self._s0 = None
        self.ray_steps = 0       # RK4 steps taken inside the cube by the last solve()
        self.last_solve_seconds = None

    @classmethod
    def legacy(cls, x, y, z, extent, **kw):
        """Old call style of the example scripts: ElectronCube(x, y, z, extent, B_on=..., ...)."""
        return cls(x, y, z, extent, **kw)

    # ---- geometry -------------------------------------------------------------------------------
    @property
    def shape(self):
        return (len(self.x), len(self.y), len(self.z))

    def _geometry(self):
        """(origin, mean spacing, rectilinear?) per xyz axis."""
        info = [_axis_spacing(a, n) for a, n in ((self.x, "x"), (self.y, "y"), (self.z, "z"))]
        return tuple(i[0] for i in info), tuple(i[1] for i in info), not all(i[2] for i in info)

    @property
    def _par(self):
        try:
            return _AXIS[self.probing_direction]
        except KeyError:
            raise ValueError("probing_direction must be 'x', 'y' or 'z'") from None

    def _meshgrid(self):
        return np.meshgrid(self.x, self.y, self.z, indexing="ij")

    XX = property(lambda self: self._meshgrid()[0])
    YY = property(lambda self: self._meshgrid()[1])
    ZZ = property(lambda self: self._meshgrid()[2])

    # ---- analytic density set-ups (particle_tracker.py:147-210), evaluated on the device ---------
    def _axes_dev(self):
        torch = _lib.torch_cuda()
        ax = [torch.as_tensor(np.asarray(a, dtype=np.float64), device="cuda") for a in (self.x, self.y, self.z)]
        return torch, ax[0][:, None, None], ax[1][None, :, None], ax[2][None, None, :]

    def _set_ne_dev(self, t):
        self._ne = t.expand(*self.shape).contiguous()

    def test_null(self):
        """Null test, an empty cube (:147-152)."""
        torch = _lib.torch_cuda()
        self._ne = torch.zeros(self.shape, dtype=torch.float64, device="cuda")

    def test_slab(self, s=1, n_e0=2e23):
        """A slab with a linear gradient in x: n_e = n_e0 * (1 + s*x/extent) (:154-165)."""
        torch, X, Y, Z = self._axes_dev()
        self._set_ne_dev(n_e0 * (1.0 + s * X / float(self.extent_x)) + 0 * Y + 0 * Z)

    def test_linear_cos(self, s1=0.1, s2=0.1, n_e0=2e23, Ly=1):
        """Linearly growing sinusoidal perturbation (:167-177)."""
        torch, X, Y, Z = self._axes_dev()
        self._set_ne_dev(n_e0 * (1.0 + s1 * X / float(self.extent_x)) * (1 + s2 * torch.cos(2 * np.pi * Y / Ly)) + 0 * Z)

    def test_exponential_cos(self, n_e0=1e24, Ly=1e-3, s=2e-3):
        """Exponentially growing sinusoidal perturbation (:179-188)."""
        torch, X, Y, Z = self._axes_dev()
        self._set_ne_dev(n_e0 * 10 ** (X / s) * (1 + torch.cos(2 * np.pi * Y / Ly)) + 0 * Z)

    def test_lens(self, n_e0=1e24, LR=1e-3):
        """Normal distribution with axis along z (:190-199)."""
        torch, X, Y, Z = self._axes_dev()
        self._set_ne_dev(n_e0 * torch.exp(-(X**2 + Y**2) / LR**2) + 0 * Z)

    def test_liner(self, n_e0=1e24, LR=1e-3):
        """Normal distribution with axis along y (:201-210)."""
        torch, X, Y, Z = self._axes_dev()
        self._set_ne_dev(n_e0 * torch.exp(-(X**2 + Z**2) / LR**2) + 0 * Y)

    def external_ne(self, ne):
        """Load externally generated grid (:212-218): numpy array, torch tensor or DeviceArray of
        shape (len(x), len(y), len(z)), density in m^-3."""
        self._ne = ne.torch if isinstance(ne, DeviceArray) else ne

    def external_B(self, B):
        """Magnetic field cube, shape (len(x), len(y), len(z), 3), Tesla (example_kitchensink.py:77)."""
        self._B = B.torch if isinstance(B, DeviceArray) else B
        self._aux = None

    def external_Te(self, Te):
        """Electron temperature cube in eV (example_kitchensink.py:75), for inverse bremsstrahlung."""
        self._Te = Te.torch if isinstance(Te, DeviceArray) else Te
        self._aux = None

    def external_Z(self, Z):
        """Ionisation state: scalar or cube (example_kitchensink.py:76)."""
        self._Z = Z.torch if isinstance(Z, DeviceArray) else Z
        self._aux = None

    def set_up_interps(self):
        """Kept for the example scripts (example_kitchensink.py:79): builds the (B, kappa) grid."""
        self._aux_grid()

    def kappa(self):
        """Inverse-bremsstrahlung energy absorption coefficient (1/m) on the cube (device tensor),
        NRL formulary: kappa[cm^-1] = 3.1e-7 Z ne^2 lnL Te^-3/2 omega^-2 (1 - ne/nc)^-1/2 with ne in
        cm^-3, Te in eV; lnL = max(2, 24 - ln(sqrt(ne)/Te)) unless ``self.coulomb_log`` is set."""
        torch = _lib.torch_cuda()
        if self._Te is None:
            raise AttributeError("inv_brems=True needs external_Te(Te)")
        ne = _lib.to_device(self._ne, torch.float64)
        Te = _lib.to_device(self._Te, torch.float64).expand(*self.shape)
        Z = self._Z if isinstance(self._Z, (int, float)) else _lib.to_device(self._Z, torch.float64)
        ne_cc = ne * 1e-6
        if self.coulomb_log is None:
            lnL = torch.clamp(24.0 - torch.log(torch.sqrt(torch.clamp(ne_cc, min=1e-30)) / Te), min=2.0)
        else:
            lnL = float(self.coulomb_log)
        ne_nc = torch.clamp(ne / self.nc, max=float(self.ne_max))
        disp = torch.sqrt(torch.clamp(1.0 - ne_nc, min=1e-6))
        return 100.0 * 3.1e-7 * Z * ne_cc**2 * lnL * Te**-1.5 / (self.omega**2 * disp)

    def _aux_grid(self):
        """(B_u, B_v, B_w, kappa) in the layout/dtype of the gradient grid, or None (phase only): one pass of
        ``tt_build_aux_grid`` over the user's cubes (no whole-cube temporaries; ``kappa()`` is the same formula as a
        diagnostic)."""
        torch = _lib.torch_cuda()
        if not (self.B_on or self.inv_brems):
            return None
        if self._aux is not None:
            return self._aux
        grid = self._require_grid()
        lib = _lib.load()

        def dev(v):                         # a user cube on the device in its own floating type (FP32 stays FP32)
            t = v.torch if isinstance(v, DeviceArray) else v
            if isinstance(t, torch.Tensor):
                t = t.cuda()
                return t if t.dtype in (torch.float32, torch.float64) else t.double()
            a_ = np.asarray(t)
            return _lib.to_device(a_, torch.float32 if a_.dtype == np.float32 else torch.float64)

        ne = dev(self._ne).contiguous()
        B = Te = Z = None
        Te_s = Z_s = 0.0
        if self.B_on:
            if self._B is None:
                raise AttributeError("B_on=True needs external_B(B)")
            B = dev(self._B)
            if tuple(B.shape) != self.shape + (3,):
                raise ValueError(f"B has shape {tuple(B.shape)}, expected {self.shape + (3,)}")
        if self.inv_brems:
            if self._Te is None:
                raise AttributeError("inv_brems=True needs external_Te(Te)")
            if isinstance(self._Te, (int, float)):
                Te_s = float(self._Te)
            else:
                Te = dev(self._Te)
                if Te.numel() == 1:
                    Te_s, Te = float(Te.item()), None
            if isinstance(self._Z, (int, float)):
                Z_s = float(self._Z)
            else:
                Z = dev(self._Z)
                if Z.numel() == 1:
                    Z_s, Z = float(Z.item()), None
        cubes = [t for t in (B, Te, Z) if t is not None]
        adt = torch.float32 if cubes and all(t.dtype == torch.float32 for t in cubes) else torch.float64
        B, Te, Z = ((None if t is None else t.to(adt)) for t in (B, Te, Z))
        Te, Z = ((None if t is None else t.expand(*self.shape).contiguous()) for t in (Te, Z))
        if B is not None:
            B = B.contiguous()
        a = torch.empty(grid.shape, dtype=grid.dtype, device="cuda")
        _lib.check(lib.tt_build_aux_grid(_lib.ptr(ne), _lib.dtype_code(ne.dtype), _lib.ptr(Te), Te_s, _lib.ptr(Z),
                                         Z_s, _lib.ptr(B), _lib.dtype_code(adt), _lib.i3(self.shape), self._par, float(self.nc),
                                         float(self.ne_max), float(self.omega),
                                         float("nan") if self.coulomb_log is None else float(self.coulomb_log),
                                         int(bool(self.inv_brems)), _lib.ptr(a), _lib.dtype_code(grid.dtype),
                                         _lib.stream_ptr()), "tt_build_aux_grid")
        self._aux = a
        self._faces_aux_key = None          # (the face coefficients of the passive fields follow)
        return a

    @property
    def ne(self):
        if self._ne is None:
            raise AttributeError("ne not set: call a test_* method or external_ne first")
        if isinstance(self._ne, np.ndarray):
            return self._ne
        return self._ne.detach().cpu().numpy()

    @ne.setter
    def ne(self, v):
        self.external_ne(v)

    # ---- gradient grid ----------------------------------------------------------------------------
    def calc_dndr(self, lwl=1053e-9, ne_max=1):
        """Generate the gradient grid (replaces the interpolators of :220-241).

        omega = 2 pi c / lwl, nc = 3.14207787e-4 omega^2, ne/nc clipped at ne_max, gradients by
        central differences (one-sided on the faces)."""
        torch = _lib.torch_cuda()
        lib = _lib.load()
        if self._ne is None:
            raise AttributeError("ne not set: call a test_* method or external_ne first")
        self.omega = 2 * np.pi * (c / lwl)
        self.nc = _NC_COEFF * self.omega**2
        self.ne_max = ne_max
        self.VerdetConst = VERDET * lwl**2          # rad / (T m^2)
        self._aux = None
        origin, spacing, rect = self._geometry()
        ne = self._ne
        if tuple(ne.shape) != self.shape:
            raise ValueError(f"ne has shape {tuple(ne.shape)}, expected {self.shape}")
        if isinstance(ne, np.ndarray) and ne.dtype not in (np.float32, np.float64):
            ne = ne.astype(np.float64)
        ne_dev = _lib.to_device(ne)
        if ne_dev.dtype not in (torch.float32, torch.float64):
            ne_dev = ne_dev.to(torch.float64)
        gdt = torch.float64 if self.dtype == "float64" else torch.float32
        par = self._par
        fa = {2: (0, 1, 2), 1: (0, 2, 1), 0: (1, 2, 0)}[par]
        n = self.shape
        gshape = (n[fa[2]], n[fa[1]], n[fa[0]], 4)
        grid = self._grid                     # rebuilt in place when the cube is only refreshed
        if grid is None or tuple(grid.shape) != gshape or grid.dtype != gdt:
            self._grid = grid = None          # release the old grid before allocating (2-17 GB)
            grid = torch.empty(gshape, dtype=gdt, device="cuda")
        self._nodes = None
        if rect:      # non-uniformly spaced axes: node coordinates go to the device (include/tt_b200.h, tt_*_axes)
            self._nodes = [torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device="cuda")
                           for a in (self.x, self.y, self.z)]
            _lib.check(lib.tt_calc_dndr_axes(_lib.ptr(ne_dev), _lib.dtype_code(ne_dev.dtype), _lib.i3(n),
                                             *(_lib.ptr(a) for a in self._nodes), par, float(self.nc), float(ne_max),
                                             _lib.ptr(grid), _lib.dtype_code(gdt), _lib.stream_ptr()),
                       "tt_calc_dndr_axes")
        else:
            _lib.check(lib.tt_calc_dndr(_lib.ptr(ne_dev), _lib.dtype_code(ne_dev.dtype), _lib.i3(n), _lib.d3(spacing),
                                        par, float(self.nc), float(ne_max), _lib.ptr(grid),
                                        _lib.dtype_code(gdt), _lib.stream_ptr()), "tt_calc_dndr")
        self._grid, self._frame, self._origin, self._spacing = grid, fa, origin, spacing
        self._faces_valid = False
        self.dndx_interp, self.dndy_interp, self.dndz_interp = (_GradInterp(self, k) for k in range(3))

    def _face_grid(self):
        """The face-coefficient grid of the current gradient grid (built on first use after calc_dndr, in place when
        the cube was only refreshed), or None when ``face_grid="auto"`` and it does not fit into free device memory."""
        torch = _lib.torch_cuda()
        lib = _lib.load()
        if self._faces is not None and self._faces_valid:
            return self._faces
        nbytes = int(lib.tt_face_grid_bytes(_lib.i3(self.shape), self._par))
        if self._faces is None or self._faces.numel() * 4 != nbytes:
            self._faces = None
            if self.face_grid == "auto" and nbytes + (8 << 30) > torch.cuda.mem_get_info()[0]:
                return None
            self._faces = torch.empty(nbytes // 4, dtype=torch.float32, device="cuda")
        _lib.check(lib.tt_build_face_grid(_lib.ptr(self._grid), _lib.i3(self.shape), _lib.d3(self._spacing), self._par,
                                          _lib.ptr(self._faces), _lib.stream_ptr()), "tt_build_face_grid")
        self._faces_valid = True
        return self._faces

    def _face_aux_grid(self, aux4):
        """The second face-coefficient grid (ne/nc, B, kappa per cell face; 80 B per face cell) for the passive quantities,
        built with the gradient faces; None when ``face_grid="auto"`` and it does not fit."""
        torch = _lib.torch_cuda()
        lib = _lib.load()
        key = (None if aux4 is None else aux4.data_ptr(), self._grid.data_ptr())
        if getattr(self, "_faces_aux", None) is not None and self._faces_aux_key == key and self._faces_valid:
            return self._faces_aux
        nbytes = int(lib.tt_face_aux_grid_bytes(_lib.i3(self.shape), self._par))
        fa = getattr(self, "_faces_aux", None)
        if fa is None or fa.numel() * 4 != nbytes:
            self._faces_aux = None
            if self.face_grid == "auto" and nbytes + (8 << 30) > torch.cuda.mem_get_info()[0]:
                return None
            self._faces_aux = torch.empty(nbytes // 4, dtype=torch.float32, device="cuda")
        _lib.check(lib.tt_build_face_aux_grid(_lib.ptr(self._grid), _lib.ptr(aux4), _lib.i3(self.shape), _lib.d3(self._spacing),
                                              self._par, _lib.ptr(self._faces_aux), _lib.stream_ptr()), "tt_build_face_aux_grid")
        self._faces_aux_key = key
        return self._faces_aux

    def _require_grid(self):
        if self._grid is None:
            raise AttributeError("gradient grid not built: call calc_dndr() first")
        return self._grid

    def _grid_component(self, axis):
        """(nx, ny, nz) FP64 numpy cube of frame component holding xyz axis `axis` (or ne/nc)."""
        g = self._require_grid()
        comp = 3 if axis == 3 else self._frame.index(axis)
        # grid dims are (w, v, u): xyz axis a sits at grid dim 2 - frame.index(a)
        dims = [2 - self._frame.index(a) for a in (0, 1, 2)]
        return g[..., comp].permute(*dims).double().cpu().numpy()

    ne_nc = property(lambda self: self._grid_component(3))
    dndx = property(lambda self: self._grid_component(0) * c**2)
    dndy = property(lambda self: self._grid_component(1) * c**2)
    dndz = property(lambda self: self._grid_component(2) * c**2)

    def dndr(self, x):
        """Gradient at the locations x (3 x N, metres) -> 3 x N, zero outside the cube (:243-256)."""
        torch = _lib.torch_cuda()
        lib = _lib.load()
        g = self._require_grid()
        host = not isinstance(x, (DeviceArray, torch.Tensor))
        xd = _lib.to_device(x, torch.float64)
        if xd.dim() != 2 or xd.shape[0] != 3:
            raise ValueError("x must have shape (3, N)")
        out = torch.empty_like(xd)
        if self._nodes is not None:
            _lib.check(lib.tt_dndr_axes(_lib.ptr(g), _lib.dtype_code(g.dtype), _lib.i3(self.shape),
                                        *(_lib.ptr(a) for a in self._nodes), self._par, _lib.ptr(xd), xd.shape[1],
                                        _lib.ptr(out), _lib.stream_ptr()), "tt_dndr_axes")
            return out.cpu().numpy() if host else DeviceArray(out)
        _lib.check(lib.tt_dndr(_lib.ptr(g), _lib.dtype_code(g.dtype), _lib.i3(self.shape), _lib.d3(self._origin),
                               _lib.d3(self._spacing), self._par, _lib.ptr(xd), xd.shape[1], _lib.ptr(out),
                               _lib.stream_ptr()), "tt_dndr")
        return out.cpu().numpy() if host else DeviceArray(out)

    # ---- beam ---------------------------------------------------------------------------------------
    def init_beam(self, Np, beam_size, divergence, *, seed=None, first_ray=0):
        """Launch rays s0 (6 x Np): uniform disc of radius beam_size on the entry face, Gaussian
        divergence (:258-310).

        Default (``seed=None``): drawn on the host from numpy's global RNG in the reference's draw
        order, i.e. bit-identical to the reference after ``np.random.seed``.  With ``seed`` the rays
        are generated on the device (Philox counter RNG, ray ids first_ray .. first_ray+Np) and stay
        in HBM."""
        par = self._par
        self.extent = (self.extent_x, self.extent_y, self.extent_z)[par]
        Np = int(Np)
        if seed is not None:
            torch = _lib.torch_cuda()
            s0 = torch.empty((6, Np), dtype=torch.float64, device="cuda")
            _lib.check(_lib.load().tt_init_beam(Np, int(first_ray), int(seed), float(beam_size), float(divergence),
                                                float(self.extent), par, _lib.ptr(s0), _lib.stream_ptr()),
                       "tt_init_beam")
            self._s0 = DeviceArray(s0)
            return
        s0 = np.zeros((6, Np))
        t = 2 * np.pi * np.random.rand(Np)
        u = np.random.rand(Np) + np.random.rand(Np)
        u[u > 1] = 2 - u[u > 1]
        phi = np.pi * np.random.rand(Np)
        chi = divergence * np.random.randn(Np)
        t1, t2 = {2: (0, 1), 1: (0, 2), 0: (1, 2)}[par]
        s0[t1] = beam_size * u * np.cos(t)
        s0[t2] = beam_size * u * np.sin(t)
        s0[par] = self.extent if par == 0 else -self.extent    # 'x' launches at +extent (:287)
        s0[3 + t1] = c * np.sin(chi) * np.cos(phi)
        s0[3 + t2] = c * np.sin(chi) * np.sin(phi)
        s0[3 + par] = c * np.cos(chi)
        self._s0 = s0

    @property
    def s0(self):
        return self._s0

    @s0.setter
    def s0(self, v):
        self._s0 = v

    # ---- solve ----------------------------------------------------------------------------------------
    def solve(self, method="RK45", *, return_status=False):
        """Trace all rays of ``self.s0`` through the cube and return rf (4 x Np: p1, angle1, p2,
        angle2 on the exit plane), as :312-331.  ``method`` is accepted for compatibility; the
        integrator is the fixed-step RK4 kernel.  Sets ``sf`` (state at t = sqrt(8) extent / c),
        ``rf``, ``ray_steps``."""
        torch = _lib.torch_cuda()
        lib = _lib.load()
        grid = self._require_grid()
        if not isinstance(method, str):      # stale call style solve(ss) of the example scripts
            self._s0 = method
        if self._s0 is None:
            raise AttributeError("s0 not set: call init_beam() first")
        if not hasattr(self, "extent"):
            self.extent = (self.extent_x, self.extent_y, self.extent_z)[self._par]
        host = isinstance(self._s0, np.ndarray)
        shape0 = tuple(self._s0.shape)
        if len(shape0) != 2 or shape0[0] not in (6, 9):
            raise ValueError("s0 must have shape (6, Np) (or (9, Np) with amplitude, phase, polarisation rows)")
        Np = shape0[1]
        use_aux = self.B_on or self.inv_brems or self.phaseshift
        nodes = self._nodes
        if nodes is not None and use_aux:
            raise NotImplementedError("B_on / inv_brems / phaseshift need uniformly spaced axes "
                                      "(the rectilinear-grid kernel integrates the trajectory only)")
        if Np == 0:                                   # empty bundle: nothing to launch
            self.rf = DeviceArray(torch.empty((4, 0), dtype=torch.float64, device="cuda"))
            self.sf = DeviceArray(torch.empty((6, 0), dtype=torch.float64, device="cuda")) if self.keep_sf else None
            self.status = DeviceArray(torch.empty(0, dtype=torch.uint8, device="cuda"))
            self._steps_dev = torch.zeros(1, dtype=torch.int64, device="cuda")
            return self.rf
        start = time()
        p = _lib.TraceParams()
        p.n_xyz[:] = self.shape
        p.origin_xyz[:] = self._origin
        p.spacing_xyz[:] = self._spacing
        p.par = self._par
        p.extent = float(self.extent)
        p.s_max = float(np.sqrt(8.0) * self.extent)
        p.steps_per_cell = self.steps_per_cell
        p.dtype = _lib.dtype_code(grid.dtype)
        p.variant = int(getattr(self, "kernel_variant", 0))
        ap = aux4 = None
        if use_aux:
            ap = _lib.AuxParams(float(self.omega), float(self.nc), float(self.VerdetConst))
            aux4 = self._aux_grid()
        faces = faces_aux = None
        applies = (nodes is None and grid.dtype == torch.float32 and self.steps_per_cell == 1 and p.variant == 0)
        if self.face_grid is True and not applies:
            raise ValueError("face_grid=True needs float32, uniformly spaced axes and steps_per_cell=1")
        if self.face_grid and applies:
            if use_aux:                         # both face grids or neither (the aux grid is built behind the gradient faces)
                had = self._faces is not None and self._faces_valid
                faces = self._face_grid()
                if faces is not None:
                    if not had:
                        self._faces_aux_key = None
                    faces_aux = self._face_aux_grid(aux4)
                    if faces_aux is None:
                        faces = None
            else:
                faces = self._face_grid()
        steps = torch.zeros(1, dtype=torch.int64, device="cuda")
        events = getattr(self, "_trace_events", None)     # optional CUDA-event timing of the kernel

        def trace_bundle(s0b, rf, sf, status, aux_out):
            """Morton sort + trace of one device bundle s0b (6, n); returns its permutation (or None)."""
            n = s0b.shape[1]
            stream = _lib.stream_ptr()
            perm = None
            if self.sort_rays and n > 1:
                need = C.c_size_t(0)
                _lib.check(lib.tt_sort_rays_workspace(n, C.byref(need)), "tt_sort_rays_workspace")
                ws = torch.empty(need.value, dtype=torch.uint8, device="cuda")
                perm = torch.empty(n, dtype=torch.int32, device="cuda")
                _lib.check(lib.tt_sort_rays(_lib.ptr(s0b), n, self._par, _lib.d3(self._origin), _lib.d3(self._spacing),
                                            _lib.i3(self.shape), _lib.ptr(perm), _lib.ptr(ws), need.value, stream),
                           "tt_sort_rays")
            if events is not None:
                e0 = torch.cuda.Event(enable_timing=True)
                e0.record()
            if nodes is not None:
                _lib.check(lib.tt_trace_axes(C.byref(p), *(_lib.ptr(a) for a in nodes), _lib.ptr(grid), _lib.ptr(s0b), n,
                                             _lib.ptr(perm), _lib.ptr(rf), _lib.ptr(sf), _lib.ptr(steps),
                                             _lib.ptr(status), stream), "tt_trace_axes")
            elif use_aux and faces_aux is not None:
                _lib.check(lib.tt_trace_faces_aux(C.byref(p), C.byref(ap), _lib.ptr(grid), _lib.ptr(aux4), _lib.ptr(faces),
                                                  _lib.ptr(faces_aux), _lib.ptr(s0b), n, _lib.ptr(perm), _lib.ptr(rf), _lib.ptr(sf),
                                                  _lib.ptr(aux_out), _lib.ptr(steps), _lib.ptr(status), stream),
                           "tt_trace_faces_aux")
            elif use_aux:
                _lib.check(lib.tt_trace_aux(C.byref(p), C.byref(ap), _lib.ptr(grid), _lib.ptr(aux4), _lib.ptr(s0b), n,
                                            _lib.ptr(perm), _lib.ptr(rf), _lib.ptr(sf), _lib.ptr(aux_out),
                                            _lib.ptr(steps), _lib.ptr(status), stream), "tt_trace_aux")
            elif faces is not None:
                _lib.check(lib.tt_trace_faces(C.byref(p), _lib.ptr(grid), _lib.ptr(faces), _lib.ptr(s0b), n, _lib.ptr(perm),
                                              _lib.ptr(rf), _lib.ptr(sf), _lib.ptr(steps), _lib.ptr(status), stream),
                           "tt_trace_faces")
            else:
                _lib.check(lib.tt_trace(C.byref(p), _lib.ptr(grid), _lib.ptr(s0b), n, _lib.ptr(perm), _lib.ptr(rf),
                                        _lib.ptr(sf), _lib.ptr(steps), _lib.ptr(status), stream), "tt_trace")
            if events is not None:
                e1 = torch.cuda.Event(enable_timing=True)
                e1.record()
                events.append((e0, e1))
            return perm

        def outputs(n):
            return (torch.empty((4, n), dtype=torch.float64, device="cuda"),
                    torch.empty((6, n), dtype=torch.float64, device="cuda") if self.keep_sf else None,
                    # always passed: the default kernel (event marching) flags rays for its second pass here
                    torch.empty(n, dtype=torch.uint8, device="cuda"),
                    torch.empty((3, n), dtype=torch.float64, device="cuda") if use_aux else None)

        # ---- host rays: H2D copies of chunk i+1 overlap the trace of chunk i (two streams) --------------------
        # Chunk sizes grow geometrically (first, g first, g^2 first, ... capped): tracing starts ~1.5 ms after the first
        # byte, every upload hides behind the trace of the previous chunk, and most rays travel in a few large chunks --
        # a chunk is a random 1/k sample of the beam, and the fewer rays share a cell column the lower the L1 reuse
        # of the trace kernel (measured on B200, 513^3 / 1e8 rays: 12.5 M-ray chunks 452.7 ms, 1.56 / 4.7 / 14 / 42 /
        # 37.7 M 441.6 ms, device-resident rays 422.5 ms).  How fast the chunks may grow depends on the upload rate
        # (0.9 ns per ray at 55 GB/s against 4.2 ns of tracing; with 8 ranks uploading at once a B200 box gives each
        # ~20 GB/s), so the rate of the FIRST chunk decides the schedule unless the caller fixed it.
        cap_user = getattr(self, "pipeline_chunk_rays", None)
        growth_user = getattr(self, "pipeline_growth", None)
        cap = max(int(cap_user or 50_000_000), 1)
        first = max(int(getattr(self, "pipeline_first_rays", min(1_562_500, max(cap // 8, 1)))), 1)
        if host and Np >= 2 * first:
            src = torch.from_numpy(np.ascontiguousarray(self._s0[:6], dtype=np.float64))
            rf, sf, status, aux_out = outputs(Np)
            perm = torch.empty(Np, dtype=torch.int32, device="cuda") if self.sort_rays else None
            main = torch.cuda.current_stream()
            copy = getattr(self, "_copy_stream", None) or torch.cuda.Stream()
            self._copy_stream = copy
            free = [None, None]                            # compute-done events per buffer
            bufs = [None, None]
            copy.wait_stream(main)
            lo, n, ci, growth, jump = 0, first, 0, None, 0
            while lo < Np:
                n = min(n, Np - lo)
                b = ci % 2
                with torch.cuda.stream(copy):
                    if free[b] is not None:
                        copy.wait_event(free[b])
                    if bufs[b] is None or bufs[b].shape[1] < n:
                        # staging buffers live in the COPY stream's pool: memory recycled there was last used by
                        # copy-stream work (ordered before this copy), never by a main-stream kernel that may
                        # still be running; the trace kernels read them on the main stream, hence record_stream
                        bufs[b] = torch.empty((6, n), dtype=torch.float64, device="cuda")
                        bufs[b].record_stream(main)
                    s0b = bufs[b] if n == bufs[b].shape[1] else bufs[b].reshape(-1)[:6 * n].view(6, n)
                    t_start = torch.cuda.Event(enable_timing=True)
                    t_start.record(copy)
                    for r in range(6):                     # (pageable source: staged by worker threads, _lib.h2d)
                        _lib.h2d(s0b[r], src[r, lo:lo + n])
                    ready = torch.cuda.Event(enable_timing=True)
                    ready.record(copy)
                main.wait_event(ready)
                rf_c, sf_c, st_c, ax_c = outputs(n)
                pc = trace_bundle(s0b, rf_c, sf_c, st_c, ax_c)
                rf[:, lo:lo + n] = rf_c
                status[lo:lo + n] = st_c
                if sf is not None:
                    sf[:, lo:lo + n] = sf_c
                if aux_out is not None:
                    aux_out[:, lo:lo + n] = ax_c
                if perm is not None:                       # global ray ids, chunk after chunk (a 1-ray chunk is not sorted)
                    perm[lo:lo + n] = (pc + lo) if pc is not None else torch.arange(lo, lo + n, dtype=torch.int32, device="cuda")
                free[b] = torch.cuda.Event()
                free[b].record(main)
                # the schedule follows the measured upload rate of the chunk just sent (its copy is finished or about to
                # be: the GPU still has the trace of this chunk queued, the host loses nothing by waiting for the event).
                # Pinned sources: the rate is the link's and does not change; pageable sources are staged by worker threads
                # and get faster with the size of a row (16 GB/s for the 12.5 MB rows of the first chunk, 48 GB/s beyond)
                if growth_user:
                    growth = float(growth_user)
                else:
                    ready.synchronize()
                    gbs = 48.0 * n / max(t_start.elapsed_time(ready), 1e-3) * 1e-6
                    if ci == 0:
                        self.last_upload_gbs = gbs
                    self.last_upload_gbs_max = max(gbs, getattr(self, "last_upload_gbs_max", 0.0) if ci else 0.0)
                    # the upload of chunk i+1 hides behind the trace of chunk i as long as
                    #   growth <= (trace time per ray) / (upload time per ray);
                    # measured on an 8-GPU box, every rank uploading at ~20 GB/s: 2.4 ns per ray up, 3.0 ns per ray
                    # traced (513 planes) -- doubling the chunk exposed 1.8 ns per ray of every growth step, 40 ms of
                    # a 300 ms solve.  Trace time: 6.1 ps per ray-step in FP32 on a B200, twice that in FP64.
                    up_ns = 48.0 / max(gbs, 1e-3)
                    tr_ns = (self.shape[self._par] - 1) * self.steps_per_cell * (6.1e-3 if grid.dtype == torch.float32 else 1.8e-2)
                    growth = min(3.0, max(1.0, 0.9 * tr_ns / up_ns))
                    jump = 0
                    if growth < 1.1:                       # upload-bound anyway: equal chunks, not too small
                        growth = 1.0
                        jump = max(n, min(6_250_000, cap))
                    if not cap_user:
                        cap = 50_000_000 if gbs >= 35.0 else (25_000_000 if gbs >= 15.0 else 12_500_000)
                self.last_pipeline_growth = growth
                lo += n
                ci += 1
                n = min(max(int(growth * n), jump), cap)
            init_aux = _lib.to_device(self._s0[6:9], torch.float64) if shape0[0] == 9 else None
        else:
            s0 = _lib.to_device(self._s0, torch.float64)
            init_aux = s0[6:9] if shape0[0] == 9 else None
            s0 = s0[:6].contiguous()
            rf, sf, status, aux_out = outputs(Np)
            perm = trace_bundle(s0, rf, sf, status, aux_out)
        if use_aux:
            if not self.phaseshift:
                aux_out[1].zero_()
            if init_aux is not None:       # rows 6-8 of a 9-row s0: initial amplitude, phase, polarisation
                aux_out[0] *= init_aux[0]
                aux_out[1] += init_aux[1]
                aux_out[2] += init_aux[2]
            self.amp, self.phase, self.pol = (DeviceArray(aux_out[i]) for i in range(3))
        self._aux_out = aux_out
        self._perm = perm
        if self.verbose:
            torch.cuda.current_stream().synchronize()
            self.last_solve_seconds = time() - start
            print("Ray trace completed in:\t", self.last_solve_seconds, "s")
        self._steps_dev = steps
        self.sf = DeviceArray(sf) if sf is not None else None
        self.rf = DeviceArray(rf)
        self.rf.perm = perm        # lets the detectors visit the rays in Morton order
        self.status = DeviceArray(status)
        return self.rf

    @property
    def ray_steps(self):
        d = getattr(self, "_steps_dev", None)
        return int(d.item()) if d is not None else 0

    @ray_steps.setter
    def ray_steps(self, v):
        self._steps_dev = None

    @property
    def Jf(self):
        """Jones vector (2 x Np complex: E_x, E_y) at the exit, as used by example_kitchensink.py:92-101:
        amplitude * exp(i phase) * (sin(pol), cos(pol)) -- the beam starts polarised along E_y, so
        atan(E_x/E_y) is the Faraday rotation."""
        a = getattr(self, "_aux_out", None)
        if a is None:
            raise AttributeError("Jf needs B_on, inv_brems or phaseshift and a solve()")
        amp, ph, pol = (t.cpu().numpy() for t in a)
        e = amp * np.exp(1j * ph)
        return np.stack([e * np.sin(pol), e * np.cos(pol)])

    def trace_ms(self):
        """Durations (ms) of the trace-kernel launches recorded since ``_trace_events = []`` was set;
        call after a synchronize."""
        return [a.elapsed_time(b) for a, b in (getattr(self, "_trace_events", None) or [])]

    def ray_at_exit(self):
        """rf from ``self.sf`` by linear back-projection to the plane axis = +extent (:333-380)."""
        torch = _lib.torch_cuda()
        if getattr(self, "sf", None) is None:
            raise AttributeError("sf was not kept: construct the cube with keep_sf=True (the default) and call solve() first; "
                                 "solve() already returns rf = ray_at_exit()")
        sf = _lib.to_device(self.sf, torch.float64)
        par = self._par
        t1, t2 = {2: (0, 1), 1: (0, 2), 0: (1, 2)}[par]
        tb = (sf[par] - float(self.extent)) / sf[3 + par]
        rf = torch.stack([sf[t1] - sf[3 + t1] * tb, torch.atan(sf[3 + t1] / sf[3 + par]),
                          sf[t2] - sf[3 + t2] * tb, torch.atan(sf[3 + t2] / sf[3 + par])])
        return DeviceArray(rf)

    def save_output_rays(self, fn=None):
        """Save rf as .npy, auto-named by date and time (:382-395)."""
        if fn is None:
            fn = "{} rays.npy".format(datetime.now().strftime("%Y-%m-%d_%H-%M-%S"))
        else:
            fn = "{}.npy".format(fn)
        with open(fn, "wb") as f:
            np.save(f, np.asarray(self.rf))

    # stale-API helper still called by the reference's example scripts (example_MPI.py:62,111)
    def clear_memory(self):
        self._ne = None


def dsdt(t, s, ElectronCube):
    """ODE right-hand side [v ; dndr(x)] on the flattened 6N state (:398-419).  Kept for API
    parity (the CUDA integrator does not call it); evaluates the gradient on the device."""
    Np = s.size // 6
    s = np.asarray(s).reshape(6, Np)
    sprime = np.zeros_like(s)
    sprime[3:6, :] = np.asarray(ElectronCube.dndr(np.ascontiguousarray(s[:3, :])))
    sprime[:3, :] = s[3:, :]
    return sprime.flatten()


def init_beam(Np, beam_size, divergence, ne_extent, probing_direction="z"):
    """Module-level variant used by the reference's (stale) example scripts
    (example_MPI.py:57, example_kitchensink.py:89): returns s0 instead of storing it."""
    ax = np.array([-ne_extent, ne_extent])
    cube = ElectronCube(ax, ax, ax, probing_direction)
    cube.init_beam(Np, beam_size, divergence)
    return cube.s0
