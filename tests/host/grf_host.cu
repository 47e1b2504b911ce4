// TEST HARNESS (not product): the half spectrum of turboGen.gaussian{1,2,3}D_FFT built on the HOST from the kernel's
// own per-mode function (csrc/grf_mode.cuh); tests/test_host_kernels.py inverse-transforms it with numpy.
#include "grf_mode.cuh"

extern "C" int host_grf_spectrum(int ndim, int N, const double* lut, const double* Wr, const double* Wi,
                                 unsigned long long seed, double* F /* interleaved re, im */) {
    using namespace tt;
    const int M = 2 * N + 1;
    const int Ma = ndim == 3 ? M : 1, Mb = ndim >= 2 ? M : 1;
    const size_t total = (size_t)Ma * Mb * (N + 1);
    const double norm = 1.0 / ((double)Ma * Mb * M);
    for (size_t i = 0; i < total; ++i) {
        const double2 o = grf_mode<double>(i, N, Ma, Mb, lut, Wr, Wi, seed, norm);
        F[2 * i] = o.x; F[2 * i + 1] = o.y;
    }
    return 0;
}
