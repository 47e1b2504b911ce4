"""Drop-in mirror of the reference's ``turboGen`` FFT generators of Gaussian random fields.

Reference: gaussian_fields/turboGen.py, gaussian3D_FFT :488-538 (Timmer & Koenig synthesis on the
odd grid M = 2N+1).  The spectrum shaping / Hermitian symmetrisation is one CUDA kernel that writes
the half spectrum in FFT order, followed by one cuFFT complex-to-real inverse transform
(csrc/grf.cu).  ``k_func`` stays an arbitrary vectorised Python callable of |k| (cycles/sample): it
is evaluated once on the host for the <= 3N^2+1 distinct values of |k| and uploaded as a table.

The cosine-mode generators gaussian{1,2,3}Dcos of the reference (:40-385) are a slower alternative
superseded by the FFT path and are out of scope.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import DeviceArray


def _sqrt_spectrum_table(N, k_func):
    M = 2 * N + 1
    q = np.arange(3 * N * N + 1, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        lut = np.sqrt(np.asarray(k_func(np.sqrt(q) / M), dtype=np.float64))
    lut = np.broadcast_to(lut, q.shape).copy()
    lut[0] = 0.0                      # F[0,0,0] = 0: zero mean (:534)
    return lut


def _gaussian_fft(ndim, N, k_func, seed, dtype, Wr, Wi, return_device):
    torch = _lib.torch_cuda()
    lib = _lib.load()
    N = int(N)
    M = 2 * N + 1
    shape = (M,) * ndim
    code = _lib.dtype_code(dtype)
    tdt = torch.float64 if code == _lib.TT_F64 else torch.float32
    lut = torch.from_numpy(_sqrt_spectrum_table(N, k_func)).cuda()
    wr = wi = None
    if seed is None or Wr is not None:
        if Wr is None:
            Wr = np.random.randn(*shape)        # draw order of the reference: Wr, then Wi
            Wi = np.random.randn(*shape)
        wr = _lib.to_device(Wr, torch.float64)
        wi = _lib.to_device(Wi, torch.float64)
        if tuple(wr.shape) != shape or tuple(wi.shape) != shape:
            raise ValueError(f"Wr and Wi must have shape {shape}")
    need = C.c_size_t(0)
    _lib.check(lib.tt_grf_nd_workspace(ndim, N, code, C.byref(need)), "tt_grf_nd_workspace")
    ws = torch.empty(need.value, dtype=torch.uint8, device="cuda")
    out = torch.empty(shape, dtype=tdt, device="cuda")
    _lib.check(lib.tt_grf_nd(ndim, N, code, _lib.ptr(lut), _lib.ptr(wr), _lib.ptr(wi), int(seed or 0), _lib.ptr(out),
                             _lib.ptr(ws), need.value, _lib.stream_ptr()), "tt_grf_nd")
    if return_device:
        return DeviceArray(out)
    return out.cpu().numpy()


def gaussian3D_FFT(N, k_func, *, seed=None, dtype="float64", Wr=None, Wi=None, return_device=False):
    """A FFT based generator for scalar Gaussian fields in 3D (:488-538).

    N: the domain is (2N+1)^3; k_func: power spectrum P(|k|).  Returns the (M, M, M) field.

    Default (``seed=None``): the two white-noise cubes are drawn on the host with
    ``np.random.randn(M, M, M)`` in the reference's order (Wr then Wi, :522-523), so that the result
    equals the reference's for the same numpy seed (to FFT rounding).  With ``seed`` the noise comes
    from the device Philox generator and nothing of size M^3 ever touches the host."""
    return _gaussian_fft(3, N, k_func, seed, dtype, Wr, Wi, return_device)


def gaussian2D_FFT(N, k_func, *, seed=None, dtype="float64", Wr=None, Wi=None, return_device=False):
    """2-D variant, domain (2N+1)^2 (:436-486)."""
    return _gaussian_fft(2, N, k_func, seed, dtype, Wr, Wi, return_device)


def gaussian1D_FFT(N, k_func, *, seed=None, dtype="float64", Wr=None, Wi=None, return_device=False):
    """1-D variant, domain 2N+1 (:388-434)."""
    return _gaussian_fft(1, N, k_func, seed, dtype, Wr, Wi, return_device)
