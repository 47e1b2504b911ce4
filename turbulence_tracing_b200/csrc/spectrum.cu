// Shell-averaged power spectrum of a 3-D scalar field on the device.  Replaces
// gaussian_fields/calculate_spectrum_3d.py:spectrum_3D_scalar (:3-59), which builds fftshift-ed complex,
// K, sort-index and filtered copies of the whole cube on the host (tens of GB at 513^3): here one
// real-to-complex cuFFT and one reduction kernel over the half spectrum (Hermitian partners counted
// twice) accumulate sum(|F|^2) and the count per |k| shell; the host divides.
#include "common.cuh"

#include <cufft.h>

namespace tt {

static constexpr int kMaxBins = 2048;

template <typename T> struct SpecC;
template <> struct SpecC<float> { typedef float2 type; };
template <> struct SpecC<double> { typedef double2 type; };

struct SpecArgs {
    int n[3];
    double val[3];      // 1 / (n * dx): numpy.fft.fftfreq computes k = i * val
    double width;       // shell width = K_max / k_bin_num
    int nbins;          // shells that are filled: k_bin_num - 1 (the reference's loop stops there, :53)
};

template <typename T>
__global__ void __launch_bounds__(256) spectrum_shell_kernel(const typename SpecC<T>::type* __restrict__ F, SpecArgs A,
                                                             double* __restrict__ sum, unsigned long long* __restrict__ cnt) {
    extern __shared__ unsigned char smem_raw[];
    double* ssum = reinterpret_cast<double*>(smem_raw);
    unsigned* scnt = reinterpret_cast<unsigned*>(ssum + A.nbins);
    for (int i = threadIdx.x; i < A.nbins; i += blockDim.x) { ssum[i] = 0.0; scnt[i] = 0u; }
    __syncthreads();
    const int nzh = A.n[2] / 2 + 1;
    const size_t total = (size_t)A.n[0] * A.n[1] * nzh;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % nzh);
        const int b = (int)((i / nzh) % A.n[1]);
        const int a = (int)(i / ((size_t)nzh * A.n[1]));
        // signed FFT frequencies (numpy.fft.fftfreq: indices above (n-1)/2 are negative)
        const int fa = a <= (A.n[0] - 1) / 2 ? a : a - A.n[0];
        const int fb = b <= (A.n[1] - 1) / 2 ? b : b - A.n[1];
        const double kx = (double)fa * A.val[0], ky = (double)fb * A.val[1], kz = (double)c * A.val[2];
        const double K = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(kx, kx), __dmul_rn(ky, ky)), __dmul_rn(kz, kz)));
        int s = (int)(K / A.width);
        // exactly the reference's comparisons  K >= (i-1)*w  and  K < i*w   (:54-55)
        while (s > 0 && K < __dmul_rn((double)s, A.width)) --s;
        while (K >= __dmul_rn((double)(s + 1), A.width)) ++s;
        if (s >= A.nbins) continue;
        const typename SpecC<T>::type v = F[i];
        const double p = (double)v.x * (double)v.x + (double)v.y * (double)v.y;
        // the conjugate partner (-k) is not stored for 0 < c < n/2 (and c = n/2 only exists once for even n)
        const bool twice = c > 0 && !(A.n[2] % 2 == 0 && c == A.n[2] / 2);
        atomicAdd(&ssum[s], twice ? 2.0 * p : p);
        atomicAdd(&scnt[s], twice ? 2u : 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < A.nbins; i += blockDim.x) {
        if (scnt[i]) { atomicAdd(&sum[i], ssum[i]); atomicAdd(&cnt[i], (unsigned long long)scnt[i]); }
    }
}

static int spec_fail(cufftResult r, const char* what) {
    set_error("cuFFT error %d in %s", (int)r, what);
    return TT_ERR_CUDA;
}
static inline size_t align256s(size_t b) { return (b + 255) & ~(size_t)255; }

static int spec_plan(const int n[3], int dtype, cufftHandle* plan, size_t* work) {
    cufftResult r = cufftCreate(plan);
    if (r != CUFFT_SUCCESS) return spec_fail(r, "cufftCreate");
    r = cufftSetAutoAllocation(*plan, 0);
    if (r == CUFFT_SUCCESS) r = cufftMakePlan3d(*plan, n[0], n[1], n[2], dtype == TT_F32 ? CUFFT_R2C : CUFFT_D2Z, work);
    if (r != CUFFT_SUCCESS) { cufftDestroy(*plan); return spec_fail(r, "cufftMakePlan3d(R2C)"); }
    return TT_OK;
}

}  // namespace tt

extern "C" int tt_spectrum3d_workspace(const int n_xyz[3], int dtype, size_t* bytes) {
    using namespace tt;
    TT_REQUIRE(n_xyz && bytes, "tt_spectrum3d_workspace: null pointer");
    TT_REQUIRE(dtype == TT_F32 || dtype == TT_F64, "tt_spectrum3d: dtype must be TT_F32 or TT_F64");
    for (int i = 0; i < 3; ++i) TT_REQUIRE(n_xyz[i] >= 2, "tt_spectrum3d: every axis needs >= 2 points");
    cufftHandle plan;
    size_t work = 0;
    int rc = spec_plan(n_xyz, dtype, &plan, &work);
    if (rc) return rc;
    cufftDestroy(plan);
    const size_t spec = (size_t)n_xyz[0] * n_xyz[1] * (n_xyz[2] / 2 + 1) * (dtype == TT_F32 ? 8 : 16);
    *bytes = align256s(spec) + align256s(work);
    return TT_OK;
}

extern "C" int tt_spectrum3d(const void* data_dev, int dtype, const int n_xyz[3], double dx, double k_max,
                             int k_bin_num, double* sum_dev, unsigned long long* count_dev, void* workspace_dev,
                             size_t workspace_bytes, tt_stream_t stream) {
    using namespace tt;
    TT_REQUIRE(data_dev && n_xyz && sum_dev && count_dev && workspace_dev, "tt_spectrum3d: null pointer");
    TT_REQUIRE(dtype == TT_F32 || dtype == TT_F64, "tt_spectrum3d: dtype must be TT_F32 or TT_F64");
    TT_REQUIRE(k_bin_num >= 2 && k_bin_num <= kMaxBins, "tt_spectrum3d: k_bin_num must be in [2, %d]", kMaxBins);
    TT_REQUIRE(dx > 0 && k_max > 0, "tt_spectrum3d: dx and k_max must be > 0");
    size_t need = 0;
    int rc = tt_spectrum3d_workspace(n_xyz, dtype, &need);
    if (rc) return rc;
    TT_REQUIRE(workspace_bytes >= need, "tt_spectrum3d: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    cufftHandle plan;
    size_t work = 0;
    rc = spec_plan(n_xyz, dtype, &plan, &work);
    if (rc) return rc;
    const size_t spec = align256s((size_t)n_xyz[0] * n_xyz[1] * (n_xyz[2] / 2 + 1) * (dtype == TT_F32 ? 8 : 16));
    cufftResult r = cufftSetWorkArea(plan, (char*)workspace_dev + spec);
    if (r == CUFFT_SUCCESS) r = cufftSetStream(plan, s);
    if (r == CUFFT_SUCCESS) {
        // out-of-place: the input field is left untouched
        if (dtype == TT_F32) r = cufftExecR2C(plan, (cufftReal*)data_dev, (cufftComplex*)workspace_dev);
        else r = cufftExecD2Z(plan, (cufftDoubleReal*)data_dev, (cufftDoubleComplex*)workspace_dev);
    }
    cufftDestroy(plan);
    if (r != CUFFT_SUCCESS) return spec_fail(r, "cuFFT R2C exec");
    SpecArgs A;
    for (int i = 0; i < 3; ++i) { A.n[i] = n_xyz[i]; A.val[i] = 1.0 / ((double)n_xyz[i] * dx); }
    A.width = k_max / (double)k_bin_num;
    A.nbins = k_bin_num - 1;
    TT_CUDA(cudaMemsetAsync(sum_dev, 0, sizeof(double) * k_bin_num, s));
    TT_CUDA(cudaMemsetAsync(count_dev, 0, sizeof(unsigned long long) * k_bin_num, s));
    const size_t smem = (size_t)A.nbins * (sizeof(double) + sizeof(unsigned));
    const int blocks = 148 * 8;
    if (dtype == TT_F32)
        spectrum_shell_kernel<float><<<blocks, 256, smem, s>>>((const float2*)workspace_dev, A, sum_dev, count_dev);
    else
        spectrum_shell_kernel<double><<<blocks, 256, smem, s>>>((const double2*)workspace_dev, A, sum_dev, count_dev);
    return launch_check("spectrum_shell_kernel");
}
