// One element of the Hermitian half spectrum of turboGen.gaussian{1,2,3}D_FFT (see grf.cu): host + device, so that
// tests/host/grf_host.cu can build the spectrum on the CPU and check the index algebra (fftshift / flip / ifftshift /
// DC = 0 / normalisation) against the reference's fields with numpy's inverse FFT.
#pragma once
#include "common.cuh"

namespace tt {

template <typename T> struct Cplx;
template <> struct Cplx<float> { typedef float2 type; };
template <> struct Cplx<double> { typedef double2 type; };

TT_HD void philox_normals(uint64_t idx, uint64_t seed, double& wr, double& wi) {
    const double two_pi = 6.283185307179586476925;
    Philox p = philox4x32_10(idx, 0x47524621ull /* 'GRF!' */, seed);
    Philox q = philox4x32_10(idx, 0x47524622ull, seed);
    wr = sqrt(-2.0 * log(u01(p.c[0], p.c[1]))) * cos(two_pi * u01(p.c[2], p.c[3]));
    wi = sqrt(-2.0 * log(u01(q.c[0], q.c[1]))) * cos(two_pi * u01(q.c[2], q.c[3]));
}

// element i = (a, b, c) of the half spectrum F[a][b][0..N] in FFT (unshifted) order
template <typename T>
TT_HD typename Cplx<T>::type grf_mode(size_t i, int N, int Ma, int Mb, const double* __restrict__ lut,
                                      const double* __restrict__ Wr, const double* __restrict__ Wi, uint64_t seed,
                                      double norm) {
    // Ma, Mb: extent of the two leading axes (M for a 3-D field; 1 for the axes a 1-D / 2-D field lacks)
    const int M = 2 * N + 1, Nh = N + 1;
    const int c = (int)(i % Nh);
    const int b = (int)((i / Nh) % Mb);
    const int a = (int)(i / ((size_t)Nh * Mb));
    const int Na = Ma == 1 ? 0 : N, Nb = Mb == 1 ? 0 : N;
    // signed frequencies of FFT-order indices
    const int fa = a <= Na ? a : a - Ma, fb = b <= Nb ? b : b - Mb, fc = c;
    // index of +k and of -k in the fftshift-ed arrays Wr, Wi (:520-526)
    const size_t jp = ((size_t)(Na + fa) * Mb + (Nb + fb)) * M + (N + fc);
    const size_t jm = ((size_t)(Na - fa) * Mb + (Nb - fb)) * M + (N - fc);
    double wrp, wip, wrm, wim;
    if (Wr) {
        wrp = Wr[jp]; wrm = Wr[jm]; wip = Wi[jp]; wim = Wi[jm];
    } else {
        philox_normals(jp, seed, wrp, wip);
        philox_normals(jm, seed, wrm, wim);
    }
    const int q = fa * fa + fb * fb + fc * fc;
    // F[0,0,0] = 0 (:534); numpy's ifftn normalisation 1/M^3 (:536) is folded into the amplitude
    const double amp = q == 0 ? 0.0 : lut[q] * norm;
    typename Cplx<T>::type o;
    o.x = (T)((wrp + wrm) * amp);
    o.y = (T)((wip - wim) * amp);
    if (q == 0) { o.x = T(0); o.y = T(0); }
    return o;
}

}  // namespace tt
