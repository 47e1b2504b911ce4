// K3 + K4: the ray integrator.  Replaces ElectronCube.solve / dsdt / dndr / ray_at_exit
// (particle_tracker.py:312-331, 398-419, 243-256, 333-380), i.e. scipy's global-step RK45 over
// three RegularGridInterpolator objects, by ONE kernel: one thread per ray, fixed-step RK4.
//
// Formulation (not a translation of the reference):
//   * non-dimensional state: position in grid-index units held as (int cell, fraction) per
//     axis, direction d = v/c, path time s = c*t.  Equations of motion
//         dX/ds = d / h,   dd/ds = g(X),   g = -1/2 grad(ne/nc)   (the reference's dnd?/c^2).
//   * plane marching: the probing-axis coordinate W is the independent variable,
//         dU/dW = (hw/hu) du/dw,  d(d)/dW = hw g/dw,  ds/dW = hw/dw,
//     so every RK4 step starts and ends exactly on a grid (sub-)plane of the probing axis and
//     never straddles the kink of the piecewise-trilinear field at a cell face in W.
//     steps_per_cell sub-planes per cell.  Stage coordinates are clamped into the cube; a ray
//     that would leave through a side face is re-stepped to the face and frozen there.
//   * rays that are steep or move backwards (d_w <= kMarchMinDw), and rays about to hit the
//     path-time cap s_max = c*T, continue in a general arc-length RK4 with the same field.
//   * outside the cube the field is exactly zero (fill_value=0.0), so rays are straight lines:
//     the prologue moves a ray launched outside to its entry point, the epilogue projects the
//     frozen exit state to the plane par-axis = +extent (ray_at_exit) and, if asked, to time T
//     (the reference's cube.sf).
// Template parameter T is the arithmetic/grid type: float (16 B corners) or double (32 B).
#include "trace_gather_ray.cuh"    // gather_ray<T, VARIANT, AUX> (+ trace_common.cuh)

namespace tt {

// (the marching loops and the per-ray body gather_ray live in trace_gather_ray.cuh: host + device)

template <typename T, int VARIANT, bool AUX = false>
__global__ void __launch_bounds__(128, sizeof(T) == 8 ? 2 : TT_TRACE_MIN_BLOCKS) trace_kernel(const typename GridT<T>::V4* __restrict__ grid,
                                                    const double* __restrict__ s0,
                                                    const uint32_t* __restrict__ perm,
                                                    double* __restrict__ rf, double* __restrict__ sf,
                                                    unsigned long long* __restrict__ ray_steps,
                                                    uint8_t* __restrict__ status, TraceArgs A, int only_flagged,
                                                    const typename GridT<T>::V4* __restrict__ aux4 = nullptr,
                                                    double* __restrict__ aux_out = nullptr, AuxArgs AX = AuxArgs()) {
    // grid-stride loop: one ray per thread in a first pass (grid covers all rays); the second pass behind
    // the event kernel runs a small grid over the status flags (781 k one-ray CTAs that exit at once
    // would cost 1.5 ms of block scheduling for 1e8 rays)
    if (only_flagged && A.any_deferred && *A.any_deferred == 0u) return;   // nothing was deferred
    if (only_flagged && A.any_deferred && A.list_cap && A.any_deferred[1] <= A.list_cap) return;   // the list kernel did them all
    const long stride = (long)gridDim.x * blockDim.x;
    const long np32 = (A.np + 31) & ~31L;             // whole warps iterate together (shuffle below)
    for (long tid = (long)blockIdx.x * blockDim.x + threadIdx.x; tid < np32; tid += stride) {
    unsigned steps = 0;
    bool mine = tid < A.np;
    long ray = 0;
    if (mine) {
        // second pass behind trace_event_kernel: only the rays it handed over, visited in ray order
        // (a coalesced scan of the flags; the Morton order only matters for the bulk of the rays)
        ray = (perm && !only_flagged) ? (long)perm[tid] : tid;
        if (only_flagged && status[ray] != TT_RAY_DEFERRED) mine = false;
    }
    if (mine) steps = gather_ray<T, VARIANT, AUX>(grid, s0, ray, rf, sf, status, A, aux4, aux_out, AX);   // trace_gather_ray.cuh
    // ---- ray-step count: warp reduce, one atomic per warp ------------------------------------
    if (ray_steps) {
        unsigned v = steps;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(ray_steps, (unsigned long long)v);
    }
    }   // grid-stride loop
}

// ElectronCube.dndr (particle_tracker.py:243-256): one point per thread, the look-up itself in trace_gather_ray.cuh
template <typename T>
__global__ void dndr_kernel(const typename GridT<T>::V4* __restrict__ grid, TraceArgs A,
                            const double* __restrict__ pos, long npts, double* __restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npts) return;
    dndr_point<T>(grid, A, pos, npts, i, out);
}

int fill_trace_args(TraceArgs& A, const int n_xyz[3], const double origin_xyz[3],
                    const double spacing_xyz[3], int par) {
    TT_REQUIRE(n_xyz && origin_xyz && spacing_xyz, "null geometry pointer");
    TT_REQUIRE(par >= 0 && par <= 2, "par must be 0, 1 or 2 (got %d)", par);
    Frame f = frame_of(par);
    for (int k = 0; k < 3; ++k) {
        A.fa[k] = f.a[k];
        A.n[k] = n_xyz[f.a[k]];
        A.o[k] = origin_xyz[f.a[k]];
        A.h[k] = spacing_xyz[f.a[k]];
        TT_REQUIRE(A.n[k] >= 2, "every axis needs >= 2 points");
        TT_REQUIRE(A.h[k] > 0, "spacing must be > 0");
    }
    A.any_deferred = nullptr;
    A.list_cap = 0;
    A.plane_elems = (long long)A.n[0] * A.n[1];
    A.hwf = (float)A.h[2]; A.ruf = (float)(A.h[2] / A.h[0]); A.rvf = (float)(A.h[2] / A.h[1]);
    return TT_OK;
}

// ---- second pass of the event-marching paths: the general (cell-cache gather) kernel over the rays flagged TT_RAY_DEFERRED.
// The flagged rays are sparse (0.2 % of a beam as wide as the cube), so a kernel that scans the flags with one ray per
// thread runs its long marching loop with one live lane per warp (measured: 33.5 ms for 2.4e5 of 1e8 rays).  They are
// first compacted into a list of ray ids (warp-aggregated atomics; order arbitrary, results are per ray), then traced
// with full warps.  Both kernels return at once when the first pass deferred nothing.
__global__ void __launch_bounds__(256)
compact_deferred_kernel(const uint8_t* __restrict__ status, long np, unsigned int* __restrict__ words,
                        uint32_t* __restrict__ list, unsigned int cap) {
    if (words[0] == 0u) return;
    const long stride = (long)gridDim.x * blockDim.x;
    const long np32 = (np + 31) & ~31L;
    for (long tid = (long)blockIdx.x * blockDim.x + threadIdx.x; tid < np32; tid += stride) {
        const bool flagged = tid < np && status[tid] == TT_RAY_DEFERRED;
        const unsigned m = __ballot_sync(0xffffffffu, flagged);
        if (m == 0u) continue;
        const int lane = threadIdx.x & 31;
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(&words[1], (unsigned)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        const unsigned slot = base + (unsigned)__popc(m & ((1u << lane) - 1u));
        if (flagged && slot < cap) list[slot] = (uint32_t)tid;
    }
}

template <typename T>
__global__ void __launch_bounds__(128, sizeof(T) == 8 ? 2 : TT_TRACE_MIN_BLOCKS)
trace_list_kernel(const typename GridT<T>::V4* __restrict__ grid, const double* __restrict__ s0, double* __restrict__ rf,
                  double* __restrict__ sf, unsigned long long* __restrict__ ray_steps, uint8_t* __restrict__ status,
                  TraceArgs A, const uint32_t* __restrict__ list) {
    if (A.any_deferred[0] == 0u) return;
    const unsigned n = A.any_deferred[1] < A.list_cap ? A.any_deferred[1] : A.list_cap;
    const unsigned n32 = (n + 31u) & ~31u;
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n32; i += stride) {
        unsigned steps = 0;
        if (i < n) steps = gather_ray<T, 0, false>(grid, s0, (long)list[i], rf, sf, status, A, nullptr, nullptr, AuxArgs());
        if (ray_steps) {
            unsigned v = steps;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) == 0 && v) atomicAdd(ray_steps, (unsigned long long)v);
        }
    }
}

int launch_trace_second_pass(int dtype, const void* grid4, const double* s0, const uint32_t* perm, double* rf, double* sf,
                             unsigned long long* ray_steps, uint8_t* status, const TraceArgs& A0, cudaStream_t s) {
    TraceArgs A = A0;
    const int block = 128;
    long blocks = (A.np + block - 1) / block;
    if (blocks > 148 * 32) blocks = 148 * 32;
    // list of deferred ray ids: up to 16 Mi entries (64 MB of stream-ordered scratch); more than that (a beam that mostly
    // misses the cube) falls through to the flag-scan kernel below, which then has enough live lanes anyway
    uint32_t* list = nullptr;
    const unsigned cap = (unsigned)(A.np < (16L << 20) ? A.np : (16L << 20));
    if (A.any_deferred && A.np < (1L << 32)) list = (uint32_t*)scratch_alloc((size_t)cap * sizeof(uint32_t), s);
    if (list) {
        A.list_cap = cap;
        compact_deferred_kernel<<<148 * 8, 256, 0, s>>>(status, A.np, A.any_deferred, list, cap);
        int rc = launch_check("compact_deferred_kernel");
        if (rc == TT_OK) {
            if (dtype == TT_F32) trace_list_kernel<float><<<148 * 8, block, 0, s>>>((const float4*)grid4, s0, rf, sf, ray_steps, status, A, list);
            else trace_list_kernel<double><<<148 * 8, block, 0, s>>>((const double4*)grid4, s0, rf, sf, ray_steps, status, A, list);
            rc = launch_check("trace_list_kernel");
        }
        cudaFreeAsync(list, s);
        if (rc) return rc;
    }
    if (dtype == TT_F32) trace_kernel<float, 0><<<(unsigned)blocks, block, 0, s>>>((const float4*)grid4, s0, perm, rf, sf, ray_steps, status, A, 1);
    else trace_kernel<double, 0><<<(unsigned)blocks, block, 0, s>>>((const double4*)grid4, s0, perm, rf, sf, ray_steps, status, A, 1);
    return launch_check("trace_kernel");
}

}  // namespace tt

extern "C" int tt_trace(const tt_trace_params* p, const void* grid4_dev, const double* s0_dev, long np,
                        const uint32_t* perm_dev, double* rf_dev, double* sf_dev,
                        unsigned long long* ray_steps_dev, uint8_t* status_dev, tt_stream_t stream) {
    using namespace tt;
    TT_REQUIRE(p && grid4_dev && s0_dev && rf_dev, "tt_trace: null pointer");
    TT_REQUIRE(np >= 0, "tt_trace: negative ray count");
    TT_REQUIRE(p->dtype == TT_F32 || p->dtype == TT_F64, "tt_trace: dtype must be TT_F32 or TT_F64");
    TT_REQUIRE(p->steps_per_cell >= 1 && p->steps_per_cell <= 1024, "tt_trace: steps_per_cell out of range");
    TT_REQUIRE(p->s_max > 0 && p->extent == p->extent, "tt_trace: s_max must be > 0");
    TT_REQUIRE(np < (1L << 32) || !perm_dev, "tt_trace: perm is 32-bit; trace in bundles of < 2^32 rays");
    TraceArgs A;
    int rc = fill_trace_args(A, p->n_xyz, p->origin_xyz, p->spacing_xyz, p->par);
    if (rc) return rc;
    A.extent = p->extent; A.s_max = p->s_max; A.spc = p->steps_per_cell; A.np = np;
    if (np == 0) return TT_OK;
    const int block = 128;
    const long blocks = (np + block - 1) / block;
    TT_REQUIRE(blocks < (1L << 31), "tt_trace: too many rays for one launch");
    cudaStream_t s = (cudaStream_t)stream;
    // variant: 0 = auto (event marching when a status buffer is given, else the cell-cache kernel),
    //          1 = 8-corner gather per stage, 2 = cell cache, 3 = event marching (needs status_dev)
    //          4 = event marching without the packed FP32x2 arithmetic (cross-check of 3 in FP32)
    int variant = p->variant;
    TT_REQUIRE(variant >= 0 && variant <= 4, "tt_trace: unknown kernel variant %d", variant);
    TT_REQUIRE(variant < 3 || status_dev, "tt_trace: event marching (variant 3/4) needs status_dev");
    if (variant == 0) variant = status_dev ? 3 : 2;
    if (variant >= 3) {
        // stream-ordered scratch words: "did the event kernel defer any ray?" / how many
        unsigned int* flag = scratch_flag(s);
        A.any_deferred = flag;
        // event marching (packed FP32x2 arithmetic for variant 3 in FP32), then the general kernel on the deferred rays only
        int rc2 = launch_trace_event(p->dtype, variant == 3, p->steps_per_cell, grid4_dev, s0_dev, perm_dev, rf_dev, sf_dev,
                                     ray_steps_dev, status_dev, A, nullptr, nullptr, AuxArgs(), s);
        if (rc2 == TT_OK) rc2 = launch_trace_second_pass(p->dtype, grid4_dev, s0_dev, perm_dev, rf_dev, sf_dev, ray_steps_dev, status_dev, A, s);
        if (flag) cudaFreeAsync(flag, s);
        return rc2;
    }
#define TT_LAUNCH(TYPE, V4T, VAR)                                                                                \
    trace_kernel<TYPE, VAR><<<(unsigned)blocks, block, 0, s>>>((const V4T*)grid4_dev, s0_dev, perm_dev, rf_dev,   \
                                                                sf_dev, ray_steps_dev, status_dev, A, 0)
    if (p->dtype == TT_F32) { if (variant == 1) TT_LAUNCH(float, float4, 1); else TT_LAUNCH(float, float4, 0); }
    else { if (variant == 1) TT_LAUNCH(double, double4, 1); else TT_LAUNCH(double, double4, 0); }
#undef TT_LAUNCH
    return launch_check("trace_kernel");
}

namespace tt {
// the general kernel with the passive quantities over the rays an event kernel flagged TT_RAY_DEFERRED (FP32 grids)
int launch_trace_aux_second_pass(const void* grid4, const void* aux4, const double* s0, const uint32_t* perm, double* rf, double* sf,
                                 double* aux_out, unsigned long long* ray_steps, uint8_t* status, const TraceArgs& A,
                                 const AuxArgs& AX, cudaStream_t s) {
    const int block = 128;
    const long blocks = (A.np + block - 1) / block;
    const long blocks2 = blocks > 148 * 32 ? 148 * 32 : blocks;
    trace_kernel<float, 1, true><<<(unsigned)blocks2, block, 0, s>>>((const float4*)grid4, s0, perm, rf, sf, ray_steps, status, A, 1,
                                                                     (const float4*)aux4, aux_out, AX);
    return launch_check("trace_kernel<aux>");
}
}  // namespace tt

extern "C" int tt_trace_aux(const tt_trace_params* p, const tt_aux_params* a, const void* grid4_dev,
                            const void* aux4_dev, const double* s0_dev, long np, const uint32_t* perm_dev,
                            double* rf_dev, double* sf_dev, double* aux_out_dev, unsigned long long* ray_steps_dev,
                            uint8_t* status_dev, tt_stream_t stream) {
    using namespace tt;
    TT_REQUIRE(p && a && grid4_dev && s0_dev && rf_dev && aux_out_dev, "tt_trace_aux: null pointer");
    TT_REQUIRE(np >= 0, "tt_trace_aux: negative ray count");
    TT_REQUIRE(p->dtype == TT_F32 || p->dtype == TT_F64, "tt_trace_aux: dtype must be TT_F32 or TT_F64");
    TT_REQUIRE(p->steps_per_cell >= 1 && p->steps_per_cell <= 1024, "tt_trace_aux: steps_per_cell out of range");
    TT_REQUIRE(p->s_max > 0, "tt_trace_aux: s_max must be > 0");
    TT_REQUIRE(np < (1L << 32) || !perm_dev, "tt_trace_aux: perm is 32-bit; trace in bundles of < 2^32 rays");
    TT_REQUIRE(a->omega > 0 && a->nc > 0, "tt_trace_aux: omega and nc must be > 0");
    TraceArgs A;
    int rc = fill_trace_args(A, p->n_xyz, p->origin_xyz, p->spacing_xyz, p->par);
    if (rc) return rc;
    A.extent = p->extent; A.s_max = p->s_max; A.spc = p->steps_per_cell; A.np = np;
    if (np == 0) return TT_OK;
    AuxArgs AX;
    AX.omega_over_c = a->omega / kC;
    AX.verdet_nc = a->verdet * a->nc;
    const int block = 128;
    const long blocks = (np + block - 1) / block;
    TT_REQUIRE(blocks < (1L << 31), "tt_trace_aux: too many rays for one launch");
    cudaStream_t s = (cudaStream_t)stream;
    TT_REQUIRE(p->variant >= 0 && p->variant <= 4, "tt_trace_aux: unknown kernel variant %d", p->variant);
    int only_flagged = 0;
    unsigned int* flag = nullptr;
    if (p->dtype == TT_F32 && status_dev && (p->variant == 0 || p->variant == 3)) {
        flag = scratch_flag(s);
        A.any_deferred = flag;
        // event marching with the passive quantities on board; the gather kernel then redoes the deferred rays
        int rc2 = launch_trace_event(TT_F32, true, p->steps_per_cell, grid4_dev, s0_dev, perm_dev, rf_dev, sf_dev,
                                     ray_steps_dev, status_dev, A, aux4_dev, aux_out_dev, AX, s);
        if (rc2) { if (flag) cudaFreeAsync(flag, s); return rc2; }
        only_flagged = 1;
    }
    const long blocks2 = only_flagged && blocks > 148 * 32 ? 148 * 32 : blocks;
    if (p->dtype == TT_F32)
        trace_kernel<float, 1, true><<<(unsigned)blocks2, block, 0, s>>>((const float4*)grid4_dev, s0_dev, perm_dev, rf_dev, sf_dev,
                                                                         ray_steps_dev, status_dev, A, only_flagged,
                                                                         (const float4*)aux4_dev, aux_out_dev, AX);
    else
        trace_kernel<double, 1, true><<<(unsigned)blocks2, block, 0, s>>>((const double4*)grid4_dev, s0_dev, perm_dev, rf_dev, sf_dev,
                                                                          ray_steps_dev, status_dev, A, 0,
                                                                          (const double4*)aux4_dev, aux_out_dev, AX);
    if (flag) cudaFreeAsync(flag, s);
    return launch_check("trace_kernel<aux>");
}

extern "C" int tt_dndr(const void* grid4_dev, int grid_dtype, const int n_xyz[3], const double origin_xyz[3],
                       const double spacing_xyz[3], int par, const double* pos_dev, long npts,
                       double* out_dev, tt_stream_t stream) {
    using namespace tt;
    TT_REQUIRE(grid4_dev && pos_dev && out_dev, "tt_dndr: null pointer");
    TT_REQUIRE(grid_dtype == TT_F32 || grid_dtype == TT_F64, "tt_dndr: dtype must be TT_F32 or TT_F64");
    TraceArgs A;
    int rc = fill_trace_args(A, n_xyz, origin_xyz, spacing_xyz, par);
    if (rc) return rc;
    A.extent = 0; A.s_max = 0; A.spc = 1; A.np = npts;
    if (npts <= 0) return TT_OK;
    const int block = 256;
    const long blocks = (npts + block - 1) / block;
    cudaStream_t s = (cudaStream_t)stream;
    if (grid_dtype == TT_F32)
        dndr_kernel<float><<<(unsigned)blocks, block, 0, s>>>((const float4*)grid4_dev, A, pos_dev, npts, out_dev);
    else
        dndr_kernel<double><<<(unsigned)blocks, block, 0, s>>>((const double4*)grid4_dev, A, pos_dev, npts, out_dev);
    return launch_check("dndr_kernel");
}
