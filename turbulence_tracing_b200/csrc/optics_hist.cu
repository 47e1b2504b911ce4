// K5 + K6: ray-transfer-matrix optics and detector binning in one pass over the rays.
// Replaces the element functions of ray_transfer_matrix.py:37-154 (each of which allocates a new
// 4 x N array), the detector programs (:208-299) and Rays.histogram (:173-195, numpy.histogram2d).
//
// One thread handles kRaysPerThread rays; the 4-vector (x, theta, y, phi) lives in registers
// while the whole element program is applied (FP64, 32 B read per ray), then the ray is binned.
// Binning is privatised per CTA: a kTile x kTile window of uint32 counters in shared memory is
// anchored at the smallest bin touched by the CTA's rays; when a CTA's rays land in a compact patch
// of the detector (spatially ordered rays, or the Morton order of the trace via perm) almost every
// increment is a shared-memory atomic and the window is flushed once with one global atomic per
// non-empty bin.  Rays outside the window fall through to a global atomic (unordered beams: ~4e10
// atomics/s on B200, still faster than a permuted gather of the rays -- see ray_transfer_matrix.py).
#include "optics_program.cuh"      // OpticsArgs, apply_program, bin_of (host + device)

namespace tt {

static constexpr int kThreads = 256;
static constexpr int kRaysPerThread = 8;
static constexpr int kTile = 64;

__global__ void __launch_bounds__(kThreads) optics_hist_kernel(const double* __restrict__ rf_in,
                                                               const uint32_t* __restrict__ perm,
                                                               const double* __restrict__ xe,
                                                               const double* __restrict__ ye,
                                                               unsigned long long* __restrict__ H,
                                                               double* __restrict__ rf_out, OpticsArgs A,
                                                               const double* __restrict__ weights,
                                                               double* __restrict__ Hw) {
    __shared__ unsigned tile[kTile * kTile];
    __shared__ int anchor[2];
    for (int i = threadIdx.x; i < kTile * kTile; i += kThreads) tile[i] = 0u;
    if (threadIdx.x < 2) anchor[threadIdx.x] = 0x7fffffff;
    __syncthreads();

    const long base = (long)blockIdx.x * (kThreads * kRaysPerThread);
    int bx[kRaysPerThread], by[kRaysPerThread];
    int mnx = 0x7fffffff, mny = 0x7fffffff;
#pragma unroll
    for (int k = 0; k < kRaysPerThread; ++k) {
        const long i = base + (long)k * kThreads + threadIdx.x;
        bx[k] = by[k] = -1;
        if (i < A.np) {
            const long ray = perm ? (long)perm[i] : i;
            double x = rf_in[ray] * A.pos_scale, th = rf_in[A.np + ray];
            double y = rf_in[2 * A.np + ray] * A.pos_scale, ph = rf_in[3 * A.np + ray];
            apply_program(A, x, th, y, ph);
            if (rf_out) {
                rf_out[ray] = x; rf_out[A.np + ray] = th; rf_out[2 * A.np + ray] = y; rf_out[3 * A.np + ray] = ph;
            }
            if (H || Hw) {
                const int ix = bin_of(x, xe, A.nbx), iy = bin_of(y, ye, A.nby);
                if (ix >= 0 && iy >= 0) {
                    bx[k] = ix; by[k] = iy;
                    mnx = min(mnx, ix); mny = min(mny, iy);
                    // weighted image (numpy.histogram2d(weights=)): FP64 atomics straight to global
                    if (Hw) atomicAdd(&Hw[(size_t)iy * A.nbx + ix], weights[ray]);
                }
            }
        }
    }
    if (!H) return;      // (uniform) rf only, or weighted image only
    // CTA-wide anchor = smallest occupied bin in x and y
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
        mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(&anchor[0], mnx); atomicMin(&anchor[1], mny); }
    __syncthreads();
    const int ax = anchor[0], ay = anchor[1];
#pragma unroll
    for (int k = 0; k < kRaysPerThread; ++k) {
        if (bx[k] < 0) continue;
        const int tx = bx[k] - ax, ty = by[k] - ay;
        if (tx < kTile && ty < kTile) atomicAdd(&tile[ty * kTile + tx], 1u);
        else atomicAdd(&H[(size_t)by[k] * A.nbx + bx[k]], 1ull);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kTile * kTile; i += kThreads) {
        const unsigned c = tile[i];
        if (c) atomicAdd(&H[(size_t)(ay + i / kTile) * A.nbx + (ax + i % kTile)], (unsigned long long)c);
    }
}

}  // namespace tt

extern "C" int tt_optics_hist_weighted(const double* rf_in_dev, long np, const uint32_t* perm_dev, double pos_scale,
                                       const tt_optic* program, int n_ops, const double* xedges_dev, int nbx,
                                       const double* yedges_dev, int nby, unsigned long long* H_dev,
                                       const double* weights_dev, double* Hw_dev, double* rf_out_dev,
                                       tt_stream_t stream) {
    using namespace tt;
    TT_REQUIRE(rf_in_dev, "tt_optics_hist: null rf_in");
    TT_REQUIRE((weights_dev == nullptr) == (Hw_dev == nullptr), "tt_optics_hist: weights and Hw go together");
    TT_REQUIRE(np >= 0, "tt_optics_hist: negative ray count");
    TT_REQUIRE(n_ops >= 0 && n_ops <= TT_MAX_OPTICS, "tt_optics_hist: program length %d exceeds TT_MAX_OPTICS", n_ops);
    TT_REQUIRE(n_ops == 0 || program, "tt_optics_hist: null program");
    TT_REQUIRE(H_dev || rf_out_dev || Hw_dev, "tt_optics_hist: nothing to compute (H, Hw and rf_out all null)");
    if (H_dev || Hw_dev) TT_REQUIRE(xedges_dev && yedges_dev && nbx >= 1 && nby >= 1, "tt_optics_hist: histogram needs edges and bin counts >= 1");
    OpticsArgs A;
    for (int i = 0; i < n_ops; ++i) {
        A.ops[i] = program[i];
        TT_REQUIRE(program[i].op >= TT_OP_DISTANCE && program[i].op <= TT_OP_KNIFE_EDGE, "tt_optics_hist: unknown op %d", program[i].op);
        if (program[i].op == TT_OP_KNIFE_EDGE)
            TT_REQUIRE(program[i].b == 1.0 || program[i].b == -1.0 || program[i].b == 2.0 || program[i].b == -2.0,
                       "tt_optics_hist: knife edge b must be +-1 (x) or +-2 (y)");
    }
    A.n_ops = n_ops; A.pos_scale = pos_scale; A.nbx = nbx; A.nby = nby; A.np = np;
    if (np == 0) return TT_OK;
    const long per = (long)kThreads * kRaysPerThread;
    const long blocks = (np + per - 1) / per;
    TT_REQUIRE(blocks < (1L << 31), "tt_optics_hist: too many rays for one launch");
    optics_hist_kernel<<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(rf_in_dev, perm_dev, xedges_dev,
                                                                                yedges_dev, H_dev, rf_out_dev, A,
                                                                                weights_dev, Hw_dev);
    return launch_check("optics_hist_kernel");
}

extern "C" int tt_optics_hist_perm(const double* rf_in_dev, long np, const uint32_t* perm_dev, double pos_scale,
                                   const tt_optic* program, int n_ops, const double* xedges_dev, int nbx,
                                   const double* yedges_dev, int nby, unsigned long long* H_dev,
                                   double* rf_out_dev, tt_stream_t stream) {
    return tt_optics_hist_weighted(rf_in_dev, np, perm_dev, pos_scale, program, n_ops, xedges_dev, nbx, yedges_dev, nby,
                                   H_dev, nullptr, nullptr, rf_out_dev, stream);
}

extern "C" int tt_optics_hist(const double* rf_in_dev, long np, double pos_scale, const tt_optic* program, int n_ops,
                              const double* xedges_dev, int nbx, const double* yedges_dev, int nby,
                              unsigned long long* H_dev, double* rf_out_dev, tt_stream_t stream) {
    return tt_optics_hist_perm(rf_in_dev, np, nullptr, pos_scale, program, n_ops, xedges_dev, nbx, yedges_dev, nby,
                               H_dev, rf_out_dev, stream);
}
