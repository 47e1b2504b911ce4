"""ORACLE -- TEST / MEASUREMENT INFRASTRUCTURE ONLY.  The reference's OWN modules, imported unmodified.

``bench.py --impl reference`` and the ``cpu_baseline`` leg of the GPU arm time the reference's own
``ElectronCube.solve`` (particle_tracker.py:312-331) fanned out with ``multiprocessing.Pool.map`` as
``example_multiprocess.py:41-51`` does.  The modules come from ``oracle/_ref/`` (``make -C oracle ref``: byte-identical
copies, git-ignored, shipped to the GPU box with the snapshot) or, in the build container, from ``/root/reference``.
``matplotlib`` is absent from this image and only used by the reference's plotting helpers: it is stubbed in
``sys.modules``.  Nothing here is imported by the product package.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(HERE, "_ref")
_LIVE = os.environ.get("TT_REFERENCE", "/root/reference")
_mods = None


def source_dirs():
    """directories holding particle_tracker.py / ray_transfer_matrix.py of the reference, or None"""
    if os.path.exists(os.path.join(_REF_DIR, "particle_tracker.py")):
        return [_REF_DIR], "oracle/_ref (unmodified copy of the reference's modules)"
    if os.path.exists(os.path.join(_LIVE, "particle_tracking", "particle_tracker.py")):
        return [os.path.join(_LIVE, "particle_tracking"), os.path.join(_LIVE, "gaussian_fields")], _LIVE
    return None, None


def available():
    return source_dirs()[0] is not None


def load():
    """-> (particle_tracker, ray_transfer_matrix, where) of the reference itself"""
    global _mods
    if _mods is None:
        dirs, where = source_dirs()
        if dirs is None:
            raise ImportError("the reference's modules are neither under oracle/_ref (make -C oracle ref) nor at " + _LIVE)
        for name in ("matplotlib", "matplotlib.pyplot"):
            sys.modules.setdefault(name, types.ModuleType(name))
        for d in dirs:
            if d not in sys.path:
                sys.path.insert(0, d)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")              # SyntaxWarning for "\l" in the reference's docstring
            import particle_tracker as pt                # noqa: E402  (the reference's, not the product's)
            import ray_transfer_matrix as rtm            # noqa: E402
        assert os.path.dirname(os.path.abspath(pt.__file__)) in [os.path.abspath(d) for d in dirs], pt.__file__
        _mods = (pt, rtm, where)
    return _mods


class NfevCounter:
    """Counts the right-hand-side evaluations of the reference's solve_ivp calls: the module attribute
    ``particle_tracker.solve_ivp`` is wrapped (the reference's code is not touched), SURVEY section 8d."""

    def __init__(self, pt):
        self.pt, self.nfev, self._orig = pt, 0, pt.solve_ivp

    def __enter__(self):
        def counted(*a, **k):
            sol = self._orig(*a, **k)
            self.nfev += int(sol.nfev)
            return sol
        self.pt.solve_ivp = counted
        return self

    def __exit__(self, *exc):
        self.pt.solve_ivp = self._orig


def quiet(fn, *a, **k):
    """the reference prints its wall time (particle_tracker.py:325)"""
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)
