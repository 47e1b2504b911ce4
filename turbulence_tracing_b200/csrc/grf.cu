// K7: Gaussian-random-field synthesis.  Replaces turboGen.gaussian3D_FFT
// (gaussian_fields/turboGen.py:488-538): instead of building K, Wr, Wi, flip(Wr), flip(Wi), W, F
// and ifftshift(F) as eight M^3 temporaries and running a complex ifftn, one kernel writes the
// HALF spectrum F[a][b][0..N] directly in FFT (unshifted) order -- the field is Hermitian by
// construction (W = Wr + flip(Wr) + i (Wi - flip(Wi)), :525-528) -- and cuFFT runs one
// complex-to-real inverse transform.  sqrt(P(|k|)) comes from a table indexed by the integer
// q = fa^2 + fb^2 + fc^2 (|k| = sqrt(q)/M), evaluated once on the host from the user's k_func.
#include "grf_mode.cuh"

#include <cufft.h>

namespace tt {

// one thread per element of the half spectrum (a, b, c), c = 0..N; the element itself: grf_mode.cuh (host + device)
template <typename T>
__global__ void __launch_bounds__(256) grf_spectrum_kernel(int N, int Ma, int Mb, const double* __restrict__ lut,
                                                           const double* __restrict__ Wr,
                                                           const double* __restrict__ Wi, uint64_t seed,
                                                           double norm, typename Cplx<T>::type* __restrict__ F) {
    const size_t total = (size_t)Ma * Mb * (N + 1);
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    F[i] = grf_mode<T>(i, N, Ma, Mb, lut, Wr, Wi, seed, norm);
}

static int cufft_fail(cufftResult r, const char* what) {
    set_error("cuFFT error %d in %s", (int)r, what);
    return TT_ERR_CUDA;
}
#define TT_FFT(call)                                              \
    do {                                                          \
        cufftResult r_ = (call);                                  \
        if (r_ != CUFFT_SUCCESS) return cufft_fail(r_, #call);    \
    } while (0)

static inline size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

static cufftResult make_plan(cufftHandle plan, int ndim, int M, int dtype, size_t* work_bytes) {
    const cufftType t = dtype == TT_F32 ? CUFFT_C2R : CUFFT_Z2D;
    if (ndim == 1) return cufftMakePlan1d(plan, M, t, 1, work_bytes);
    if (ndim == 2) return cufftMakePlan2d(plan, M, M, t, work_bytes);
    return cufftMakePlan3d(plan, M, M, M, t, work_bytes);
}

static int plan_size(int ndim, int N, int dtype, size_t* spectrum_bytes, size_t* work_bytes) {
    const int M = 2 * N + 1;
    size_t lead = ndim == 3 ? (size_t)M * M : (ndim == 2 ? (size_t)M : 1);
    *spectrum_bytes = align256(lead * (N + 1) * (dtype == TT_F32 ? 8 : 16));
    cufftHandle plan;
    TT_FFT(cufftCreate(&plan));
    cufftResult r = cufftSetAutoAllocation(plan, 0);
    if (r == CUFFT_SUCCESS) r = make_plan(plan, ndim, M, dtype, work_bytes);
    cufftDestroy(plan);
    if (r != CUFFT_SUCCESS) return cufft_fail(r, "cufftMakePlan(size query)");
    return TT_OK;
}

}  // namespace tt

extern "C" int tt_grf_nd_workspace(int ndim, int N, int dtype, size_t* bytes) {
    using namespace tt;
    TT_REQUIRE(bytes, "tt_grf_workspace: null pointer");
    TT_REQUIRE(ndim >= 1 && ndim <= 3, "tt_grf: ndim must be 1, 2 or 3");
    TT_REQUIRE(N >= 1 && N <= 2047, "tt_grf: N out of range");
    TT_REQUIRE(dtype == TT_F32 || dtype == TT_F64, "tt_grf: dtype must be TT_F32 or TT_F64");
    size_t spec = 0, work = 0;
    int rc = plan_size(ndim, N, dtype, &spec, &work);
    if (rc) return rc;
    *bytes = spec + align256(work);
    return TT_OK;
}

extern "C" int tt_grf_nd(int ndim, int N, int dtype, const double* sqrtP_lut_dev, const double* Wr_dev,
                         const double* Wi_dev, uint64_t seed, void* out_dev, void* workspace_dev,
                         size_t workspace_bytes, tt_stream_t stream) {
    using namespace tt;
    TT_REQUIRE(sqrtP_lut_dev && out_dev && workspace_dev, "tt_grf: null pointer");
    TT_REQUIRE((Wr_dev == nullptr) == (Wi_dev == nullptr), "tt_grf: give both Wr and Wi or neither");
    TT_REQUIRE(ndim >= 1 && ndim <= 3, "tt_grf: ndim must be 1, 2 or 3");
    TT_REQUIRE(N >= 1 && N <= 2047, "tt_grf: N out of range");
    TT_REQUIRE(dtype == TT_F32 || dtype == TT_F64, "tt_grf: dtype must be TT_F32 or TT_F64");
    size_t spec = 0, work = 0;
    int rc = plan_size(ndim, N, dtype, &spec, &work);
    if (rc) return rc;
    TT_REQUIRE(workspace_bytes >= spec + align256(work), "tt_grf: workspace too small");
    const int M = 2 * N + 1;
    const int Ma = ndim == 3 ? M : 1, Mb = ndim >= 2 ? M : 1;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t total = (size_t)Ma * Mb * (N + 1);
    const unsigned blocks = (unsigned)((total + 255) / 256);
    const double norm = 1.0 / ((double)Ma * Mb * M);
    if (dtype == TT_F32)
        grf_spectrum_kernel<float><<<blocks, 256, 0, s>>>(N, Ma, Mb, sqrtP_lut_dev, Wr_dev, Wi_dev, seed, norm, (float2*)workspace_dev);
    else
        grf_spectrum_kernel<double><<<blocks, 256, 0, s>>>(N, Ma, Mb, sqrtP_lut_dev, Wr_dev, Wi_dev, seed, norm, (double2*)workspace_dev);
    rc = launch_check("grf_spectrum_kernel");
    if (rc) return rc;

    cufftHandle plan;
    TT_FFT(cufftCreate(&plan));
    size_t ws = 0;
    cufftResult r = cufftSetAutoAllocation(plan, 0);
    if (r == CUFFT_SUCCESS) r = make_plan(plan, ndim, M, dtype, &ws);
    if (r == CUFFT_SUCCESS) r = cufftSetWorkArea(plan, (char*)workspace_dev + spec);
    if (r == CUFFT_SUCCESS) r = cufftSetStream(plan, s);
    if (r == CUFFT_SUCCESS) {
        if (dtype == TT_F32) r = cufftExecC2R(plan, (cufftComplex*)workspace_dev, (cufftReal*)out_dev);
        else r = cufftExecZ2D(plan, (cufftDoubleComplex*)workspace_dev, (cufftDoubleReal*)out_dev);
    }
    cufftDestroy(plan);
    if (r != CUFFT_SUCCESS) return cufft_fail(r, "cuFFT C2R plan/exec");
    return TT_OK;
}

extern "C" int tt_grf_workspace(int N, int dtype, size_t* bytes) { return tt_grf_nd_workspace(3, N, dtype, bytes); }

extern "C" int tt_grf3d(int N, int dtype, const double* sqrtP_lut_dev, const double* Wr_dev, const double* Wi_dev,
                        uint64_t seed, void* out_dev, void* workspace_dev, size_t workspace_bytes,
                        tt_stream_t stream) {
    return tt_grf_nd(3, N, dtype, sqrtP_lut_dev, Wr_dev, Wi_dev, seed, out_dev, workspace_dev, workspace_bytes, stream);
}
