// tt_face_aux_grid_bytes / tt_build_face_aux_grid / tt_trace_faces_aux: event marching over the face-coefficient grids with
// the passive quantities (phase, Faraday rotation, absorption) on board (trace_face_aux_ray.cuh) -- BASELINE configs[3]
// in FP32 at 1 step per cell.
#include "trace_face_aux_ray.cuh"

#pragma nv_diag_suppress 550

namespace tt {

#ifndef TT_FACE_AUX_BLOCK
#define TT_FACE_AUX_BLOCK 128
#endif
#ifndef TT_FACE_AUX_MIN_BLOCKS
#define TT_FACE_AUX_MIN_BLOCKS 3
#endif

// one thread per face cell, u fastest: reads 4 corners of each node grid (L1 serves the overlap), writes 80 contiguous bytes
__global__ void __launch_bounds__(256)
face_aux_grid_kernel(const float4* __restrict__ grid, const float4* __restrict__ aux, float4* __restrict__ facesA, int nu, int nv,
                     int nw, double su, double sv) {
    const int nuc = nu - 1, nvc = nv - 1;
    const long long plane = (long long)nu * nv;
    const long long total = (long long)nuc * nvc * (nw + 1);        // + the spare plane: face nw-1 once more
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cu = (int)(i % nuc);
        const long long r = i / nuc;
        const int cv = (int)(r % nvc);
        int k = (int)(r / nvc);
        if (k > nw - 1) k = nw - 1;
        float4 out[5];
        face_aux_cell(grid, aux, nu, plane, cu, cv, k, su, sv, out);
        float4* o = facesA + 5 * i;
#pragma unroll
        for (int m = 0; m < 5; ++m) o[m] = out[m];
    }
}

template <bool TRACK_S>
__global__ void __launch_bounds__(TT_FACE_AUX_BLOCK, TT_FACE_AUX_MIN_BLOCKS)
trace_face_aux_kernel_f32x2(const float4* __restrict__ faces, const float4* __restrict__ facesA, const double* __restrict__ s0,
                            const uint32_t* __restrict__ perm, double* __restrict__ rf, double* __restrict__ sf,
                            double* __restrict__ aux_out, unsigned long long* __restrict__ ray_steps, uint8_t* __restrict__ status,
                            TraceArgs A, FaceArgs FA, AuxArgs AX) {
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned steps = 0;
    if (tid < A.np) {
        const long ray = perm ? (long)perm[tid] : tid;
        bool deferred = false;
        steps = face_aux_ray_f32x2<TRACK_S>(faces, facesA, s0, ray, rf, sf, aux_out, status, A, FA, AX, deferred);
        if (deferred && A.any_deferred) *A.any_deferred = 1u;       // (benign race: everybody stores 1)
    }
    if (ray_steps) {
        unsigned v = steps;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(ray_steps, (unsigned long long)v);
    }
}

// trace.cu: the general kernel with the passive quantities over the rays flagged TT_RAY_DEFERRED
int launch_trace_aux_second_pass(const void* grid4, const void* aux4, const double* s0, const uint32_t* perm, double* rf, double* sf,
                                 double* aux_out, unsigned long long* ray_steps, uint8_t* status, const TraceArgs& A,
                                 const AuxArgs& AX, cudaStream_t s);

}  // namespace tt

extern "C" size_t tt_face_aux_grid_bytes(const int n_xyz[3], int par) {
    if (!n_xyz || par < 0 || par > 2 || n_xyz[0] < 2 || n_xyz[1] < 2 || n_xyz[2] < 2) return 0;
    const tt::Frame f = tt::frame_of(par);
    return 80ull * (size_t)(n_xyz[f.a[0]] - 1) * (size_t)(n_xyz[f.a[1]] - 1) * ((size_t)n_xyz[f.a[2]] + 1);
}

extern "C" int tt_build_face_aux_grid(const void* grid4_dev, const void* aux4_dev, const int n_xyz[3], const double spacing_xyz[3],
                                      int par, void* faces_aux_dev, tt_stream_t stream) {
    using namespace tt;
    TT_REQUIRE(grid4_dev && faces_aux_dev && n_xyz && spacing_xyz, "tt_build_face_aux_grid: null pointer");
    TT_REQUIRE(par >= 0 && par <= 2, "tt_build_face_aux_grid: par must be 0, 1 or 2 (got %d)", par);
    const Frame f = frame_of(par);
    int n[3];
    double h[3];
    for (int k = 0; k < 3; ++k) {
        n[k] = n_xyz[f.a[k]]; h[k] = spacing_xyz[f.a[k]];
        TT_REQUIRE(n[k] >= 2, "tt_build_face_aux_grid: every axis needs >= 2 points");
        TT_REQUIRE(h[k] > 0, "tt_build_face_aux_grid: spacing must be > 0");
    }
    double su, sv;
    face_aux_scales(h, su, sv);
    const long long total = (long long)(n[0] - 1) * (n[1] - 1) * (n[2] + 1);
    const int block = 256;
    long long blocks = (total + block - 1) / block;
    if (blocks > 148LL * 64) blocks = 148LL * 64;
    face_aux_grid_kernel<<<(unsigned)blocks, block, 0, (cudaStream_t)stream>>>((const float4*)grid4_dev, (const float4*)aux4_dev,
                                                                              (float4*)faces_aux_dev, n[0], n[1], n[2], su, sv);
    return launch_check("face_aux_grid_kernel");
}

extern "C" int tt_trace_faces_aux(const tt_trace_params* p, const tt_aux_params* a, const void* grid4_dev, const void* aux4_dev,
                                  const void* faces_dev, const void* faces_aux_dev, const double* s0_dev, long np,
                                  const uint32_t* perm_dev, double* rf_dev, double* sf_dev, double* aux_out_dev,
                                  unsigned long long* ray_steps_dev, uint8_t* status_dev, tt_stream_t stream) {
    using namespace tt;
    TT_REQUIRE(p && a && grid4_dev && faces_dev && faces_aux_dev && s0_dev && rf_dev && aux_out_dev && status_dev,
               "tt_trace_faces_aux: null pointer");
    TT_REQUIRE(np >= 0, "tt_trace_faces_aux: negative ray count");
    TT_REQUIRE(p->dtype == TT_F32, "tt_trace_faces_aux: the face-coefficient grids are FP32 (use tt_trace_aux for TT_F64)");
    TT_REQUIRE(p->steps_per_cell == 1, "tt_trace_faces_aux: 1 step per cell (use tt_trace_aux for more)");
    TT_REQUIRE(p->s_max > 0 && p->extent == p->extent, "tt_trace_faces_aux: s_max must be > 0");
    TT_REQUIRE(np < (1L << 32) || !perm_dev, "tt_trace_faces_aux: perm is 32-bit; trace in bundles of < 2^32 rays");
    TT_REQUIRE(a->omega > 0 && a->nc > 0, "tt_trace_faces_aux: omega and nc must be > 0");
    TraceArgs A;
    int rc = fill_trace_args(A, p->n_xyz, p->origin_xyz, p->spacing_xyz, p->par);
    if (rc) return rc;
    A.extent = p->extent; A.s_max = p->s_max; A.spc = 1; A.np = np;
    if (np == 0) return TT_OK;
    AuxArgs AX;
    AX.omega_over_c = a->omega / kC;
    AX.verdet_nc = a->verdet * a->nc;
    FaceArgs FA;
    fill_face_args(FA, A);
    const int block = TT_FACE_AUX_BLOCK;
    const long blocks = (np + block - 1) / block;
    TT_REQUIRE(blocks < (1L << 31), "tt_trace_faces_aux: too many rays for one launch");
    cudaStream_t s = (cudaStream_t)stream;
    unsigned int* flag = scratch_flag(s);
    A.any_deferred = flag;
    if (sf_dev) trace_face_aux_kernel_f32x2<true><<<(unsigned)blocks, block, 0, s>>>((const float4*)faces_dev, (const float4*)faces_aux_dev,
                                                                                   s0_dev, perm_dev, rf_dev, sf_dev, aux_out_dev,
                                                                                   ray_steps_dev, status_dev, A, FA, AX);
    else trace_face_aux_kernel_f32x2<false><<<(unsigned)blocks, block, 0, s>>>((const float4*)faces_dev, (const float4*)faces_aux_dev,
                                                                             s0_dev, perm_dev, rf_dev, sf_dev, aux_out_dev,
                                                                             ray_steps_dev, status_dev, A, FA, AX);
    rc = launch_check("trace_face_aux_kernel_f32x2");
    if (rc == TT_OK) rc = launch_trace_aux_second_pass(grid4_dev, aux4_dev, s0_dev, perm_dev, rf_dev, sf_dev, aux_out_dev,
                                                       ray_steps_dev, status_dev, A, AX, s);
    if (flag) cudaFreeAsync(flag, s);
    return rc;
}
