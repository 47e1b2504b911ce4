"""Drop-in mirror of the reference's ``calculate_spectrum_3d.spectrum_3D_scalar``
(gaussian_fields/calculate_spectrum_3d.py:3-59): shell-averaged power spectrum of a 3-D scalar field,
computed on the device (one real-to-complex cuFFT + one shell reduction, csrc/spectrum.cu) so that the
513^3 / 1025^3 cubes of the benchmark configurations can be checked without tens of GB of host
temporaries.  The 2-D variant of the reference (:62-118) is out of scope."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def spectrum_3D_scalar(data, dx, k_bin_num=100):
    """data: (Mx, My, Mz) real field (numpy / torch / DeviceArray, float32 or float64); dx: grid spacing;
    k_bin_num: number of |k| shells.  Returns (k_bins_weighted, spect3D) exactly as the reference:
    volume-weighted shell centres and the MEAN power per shell; the last shell is left at 0 and an empty
    shell is NaN, as with the reference's loop and ``mean()`` (:53-57)."""
    torch = _lib.torch_cuda()
    lib = _lib.load()
    t = _lib.to_device(data)
    if t.dim() != 3:
        raise ValueError("data must be a 3-D array")
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float64)
    n = tuple(int(v) for v in t.shape)
    k_bin_num = int(k_bin_num)
    # K.max() as the reference finds it: fftfreq per axis, corner of the cube
    kmax = float(np.sqrt(sum(np.abs(np.fft.fftfreq(m, dx)).max() ** 2 for m in n)))
    need = C.c_size_t(0)
    code = _lib.dtype_code(t.dtype)
    _lib.check(lib.tt_spectrum3d_workspace(_lib.i3(n), code, C.byref(need)), "tt_spectrum3d_workspace")
    ws = torch.empty(need.value, dtype=torch.uint8, device="cuda")
    ssum = torch.empty(k_bin_num, dtype=torch.float64, device="cuda")
    cnt = torch.empty(k_bin_num, dtype=torch.int64, device="cuda")
    _lib.check(lib.tt_spectrum3d(_lib.ptr(t), code, _lib.i3(n), float(dx), kmax, k_bin_num, _lib.ptr(ssum),
                                 _lib.ptr(cnt), _lib.ptr(ws), need.value, _lib.stream_ptr()), "tt_spectrum3d")
    k_bin_width = kmax / k_bin_num
    k_bins = k_bin_width * np.arange(0, k_bin_num + 1)
    k_bins_weighted = (0.5 * (k_bins[:-1] ** 3 + k_bins[1:] ** 3)) ** (1 / 3)
    s, c = ssum.cpu().numpy(), cnt.cpu().numpy()
    with np.errstate(invalid="ignore", divide="ignore"):
        spect = np.where(c > 0, s / c, np.nan)
    spect[k_bin_num - 1] = 0.0
    return k_bins_weighted, spect
