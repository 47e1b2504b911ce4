"""ctypes binding of libtt_b200.so (include/tt_b200.h) + small device-array helpers.

There is deliberately NO CPU fallback: if the shared library cannot be built/loaded, or no CUDA
device is present, every compute entry point raises.  torch is used only for device memory,
streams and (elsewhere) torch.distributed.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
# TT_B200_LIB: another build of the same library (A/B runs of kernel variants, scripts/build_variants.py)
LIB_PATH = os.environ.get("TT_B200_LIB") or os.path.join(PKG, "libtt_b200.so")

TT_F32, TT_F64 = 0, 1
TT_OK = 0
RAY_EXIT_FACE, RAY_EXIT_SIDE, RAY_TIME_CAP, RAY_MISSED, RAY_GENERAL = 1, 2, 4, 8, 16
OP_DISTANCE, OP_LENS, OP_CIRC_APERTURE, OP_CIRC_STOP, OP_ANNULAR_STOP, OP_RECT_APERTURE, OP_KNIFE_EDGE = range(7)
MAX_OPTICS = 64


class TTError(RuntimeError):
    pass


class TraceParams(C.Structure):
    _fields_ = [
        ("n_xyz", C.c_int * 3),
        ("origin_xyz", C.c_double * 3),
        ("spacing_xyz", C.c_double * 3),
        ("par", C.c_int),
        ("extent", C.c_double),
        ("s_max", C.c_double),
        ("steps_per_cell", C.c_int),
        ("dtype", C.c_int),
        ("variant", C.c_int),
    ]


class AuxParams(C.Structure):
    _fields_ = [("omega", C.c_double), ("nc", C.c_double), ("verdet", C.c_double)]


class Optic(C.Structure):
    _fields_ = [("op", C.c_int), ("pad_", C.c_int), ("a", C.c_double), ("b", C.c_double)]


_I3 = C.c_int * 3
_D3 = C.c_double * 3
_vp, _i, _l, _d, _u64, _sz = C.c_void_p, C.c_int, C.c_long, C.c_double, C.c_uint64, C.c_size_t

# name -> (restype, argtypes); must list every symbol declared in include/tt_b200.h
PROTOTYPES = {
    "tt_abi_version": (_i, []),
    "tt_last_error": (C.c_char_p, []),
    "tt_device_count": (_i, []),
    "tt_launch_count": (C.c_ulonglong, []),
    "tt_calc_dndr": (_i, [_vp, _i, C.POINTER(_I3), C.POINTER(_D3), _i, _d, _d, _vp, _i, _vp]),
    "tt_dndr": (_i, [_vp, _i, C.POINTER(_I3), C.POINTER(_D3), C.POINTER(_D3), _i, _vp, _l, _vp, _vp]),
    "tt_init_beam": (_i, [_l, _l, _u64, _d, _d, _d, _i, _vp, _vp]),
    "tt_sort_rays_workspace": (_i, [_l, C.POINTER(_sz)]),
    "tt_sort_rays": (_i, [_vp, _l, _i, C.POINTER(_D3), C.POINTER(_D3), C.POINTER(_I3), _vp, _vp, _sz, _vp]),
    "tt_trace": (_i, [C.POINTER(TraceParams), _vp, _vp, _l, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tt_face_grid_bytes": (_sz, [C.POINTER(_I3), _i]),
    "tt_build_face_grid": (_i, [_vp, C.POINTER(_I3), C.POINTER(_D3), _i, _vp, _vp]),
    "tt_trace_faces": (_i, [C.POINTER(TraceParams), _vp, _vp, _vp, _l, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tt_trace_aux": (_i, [C.POINTER(TraceParams), C.POINTER(AuxParams), _vp, _vp, _vp, _l, _vp, _vp, _vp, _vp, _vp,
                          _vp, _vp]),
    "tt_face_aux_grid_bytes": (_sz, [C.POINTER(_I3), _i]),
    "tt_build_face_aux_grid": (_i, [_vp, _vp, C.POINTER(_I3), C.POINTER(_D3), _i, _vp, _vp]),
    "tt_trace_faces_aux": (_i, [C.POINTER(TraceParams), C.POINTER(AuxParams), _vp, _vp, _vp, _vp, _vp, _l, _vp, _vp, _vp, _vp,
                                _vp, _vp, _vp]),
    "tt_h2d_pageable": (_i, [_vp, _vp, _sz, _vp]),
    "tt_build_aux_grid": (_i, [_vp, _i, _vp, _d, _vp, _d, _vp, _i, C.POINTER(_I3), _i, _d, _d, _d, _d, _i, _vp, _i, _vp]),
    "tt_calc_dndr_axes": (_i, [_vp, _i, C.POINTER(_I3), _vp, _vp, _vp, _i, _d, _d, _vp, _i, _vp]),
    "tt_trace_axes": (_i, [C.POINTER(TraceParams), _vp, _vp, _vp, _vp, _vp, _l, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tt_dndr_axes": (_i, [_vp, _i, C.POINTER(_I3), _vp, _vp, _vp, _i, _vp, _l, _vp, _vp]),
    "tt_optics_hist": (_i, [_vp, _l, _d, C.POINTER(Optic), _i, _vp, _i, _vp, _i, _vp, _vp, _vp]),
    "tt_optics_hist_perm": (_i, [_vp, _l, _vp, _d, C.POINTER(Optic), _i, _vp, _i, _vp, _i, _vp, _vp, _vp]),
    "tt_optics_hist_weighted": (_i, [_vp, _l, _vp, _d, C.POINTER(Optic), _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "tt_grf_workspace": (_i, [_i, _i, C.POINTER(_sz)]),
    "tt_grf3d": (_i, [_i, _i, _vp, _vp, _vp, _u64, _vp, _vp, _sz, _vp]),
    "tt_grf_nd_workspace": (_i, [_i, _i, _i, C.POINTER(_sz)]),
    "tt_grf_nd": (_i, [_i, _i, _i, _vp, _vp, _vp, _u64, _vp, _vp, _sz, _vp]),
    "tt_spectrum3d_workspace": (_i, [C.POINTER(_I3), _i, C.POINTER(_sz)]),
    "tt_spectrum3d": (_i, [_vp, _i, C.POINTER(_I3), _d, _d, _i, _vp, _vp, _vp, _sz, _vp]),
    "tt_solve_host": (_i, [_vp, C.POINTER(_I3), C.POINTER(_D3), C.POINTER(_D3), _i, _d, _d, _d, _i, _i, _vp, _l,
                           _vp, _vp, C.POINTER(C.c_ulonglong)]),
}

_lib = None


def lib_path() -> str:
    """path of the shared library load() uses (TT_B200_LIB overrides the in-tree build)"""
    return LIB_PATH


def load(build_if_missing: bool = True):
    """Load libtt_b200.so (building it with nvcc when absent).  Raises TTError otherwise."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise TTError(f"{LIB_PATH} is missing (run `python -m turbulence_tracing_b200.build`)")
        from . import build as _build
        _build.build()
    try:
        lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    except OSError as e:  # loud: there is no other implementation to fall back to
        raise TTError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.tt_abi_version() != 1:
        raise TTError("libtt_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != TT_OK:
        msg = load().tt_last_error().decode(errors="replace")
        raise TTError(f"{what or 'libtt_b200'} failed (code {rc}): {msg}")


def torch_cuda():
    """torch with a usable CUDA device, or a loud error (no CPU fallback)."""
    import torch
    if not torch.cuda.is_available():
        raise TTError("turbulence_tracing_b200 needs a CUDA device (B200); there is no CPU fallback")
    return torch


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """device pointer of a torch tensor (or None)"""
    return None if t is None else C.c_void_p(t.data_ptr())


def i3(v):
    return C.byref(_I3(*[int(a) for a in v]))


def d3(v):
    return C.byref(_D3(*[float(a) for a in v]))


def dtype_code(dtype) -> int:
    s = str(dtype).replace("torch.", "")
    if s in ("float32", "f32", "single"):
        return TT_F32
    if s in ("float64", "f64", "double"):
        return TT_F64
    raise ValueError(f"dtype must be float32 or float64, got {dtype!r}")


class DeviceArray:
    """numpy-compatible handle of a CUDA tensor: converts to host lazily (``np.asarray(a)``,
    indexing, arithmetic) so that large results (4 x 1e8 rays) can stay in HBM while code written
    for the reference's numpy arrays keeps working.  ``.torch`` is the device tensor."""

    __array_priority__ = 100
    perm = None      # optional Morton permutation of the rays (device int32), set by ElectronCube.solve

    def __init__(self, tensor):
        self._t = tensor
        self._np = None

    @property
    def torch(self):
        return self._t

    def numpy(self):
        if self._np is None:
            self._np = self._t.detach().cpu().numpy()
        return self._np

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None and a.dtype != dtype else a

    shape = property(lambda self: tuple(self._t.shape))
    dtype = property(lambda self: self.numpy().dtype if self._np is not None else np.dtype(str(self._t.dtype).replace("torch.", "")))
    ndim = property(lambda self: self._t.dim())
    size = property(lambda self: self._t.numel())

    def __len__(self):
        return self._t.shape[0]

    def __getitem__(self, k):
        return self.numpy()[k]

    def __setitem__(self, k, v):
        import torch
        self.numpy()[k] = v
        self._t.copy_(torch.from_numpy(self._np))

    def __getattr__(self, name):           # anything else: behave like the host array
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.numpy(), name)

    def __repr__(self):
        return f"DeviceArray(shape={self.shape}, device={self._t.device})"


def _binop(name):
    def f(self, other):
        other = other.numpy() if isinstance(other, DeviceArray) else other
        return getattr(self.numpy(), name)(other)
    return f


for _n in ("add", "sub", "mul", "truediv", "pow", "radd", "rsub", "rmul", "rtruediv", "lt", "le", "gt", "ge",
           "eq", "ne", "neg", "matmul", "rmatmul"):
    if _n == "neg":
        setattr(DeviceArray, "__neg__", lambda self: -self.numpy())
    else:
        setattr(DeviceArray, f"__{_n}__", _binop(f"__{_n}__"))


def to_device(a, dtype=None):
    """numpy / torch / DeviceArray -> contiguous CUDA tensor (pinned staging for big host arrays)."""
    torch = torch_cuda()
    if isinstance(a, DeviceArray):
        t = a.torch
    elif isinstance(a, torch.Tensor):
        t = a
    else:
        t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if not t.is_cuda:
        t = h2d(None, t.contiguous())
    return t.contiguous()


H2D_PAGEABLE_MIN = 8 << 20        # bytes from which a pageable source goes through tt_h2d_pageable


def h2d(dst, src):
    """Copy the contiguous CPU tensor ``src`` into the contiguous CUDA tensor ``dst`` (allocated here if None) on the
    current stream.  Pinned sources: one asynchronous copy.  Big pageable sources (a drop-in caller's numpy arrays):
    ``tt_h2d_pageable`` -- worker threads stage them through pinned buffers, several times the rate of the driver's own
    single-threaded staging.  Returns ``dst``."""
    torch = torch_cuda()
    if dst is None:
        dst = torch.empty(src.shape, dtype=src.dtype, device="cuda")
    nbytes = src.numel() * src.element_size()
    if nbytes < H2D_PAGEABLE_MIN or src.is_pinned() or not src.is_contiguous() or not dst.is_contiguous():
        dst.copy_(src, non_blocking=True)
    else:
        check(load().tt_h2d_pageable(dst.data_ptr(), src.data_ptr(), nbytes, stream_ptr()), "tt_h2d_pageable")
    return dst
