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
//
// optics_hist_smem16_kernel (round 2): when the whole image fits ONE CTA's shared memory as 16-bit counters (88 408 bins =
// 177 KB at the default binning) every ray is binned with a shared-memory atomic whatever order the rays arrive in, each
// ray is read once, and 1e8 scattered global atomics become 148 coalesced flushes of the image.  Two counters share a
// 32-bit word; a counter never carries into its neighbour: the add that takes a counter from 0x7FFF to 0x8000 (seen in the
// value atomicAdd returns) makes its thread move 0x8000 counts to the global image, and a CTA-wide barrier per batch of
// kSmemThreads x kSmemUnroll = 2048 rays bounds what the other threads can add in between (0x7FFF + 2048 < 0xFFFF; the
// subtraction lands before the barrier, so every counter is <= 0x7FFF at every barrier).  The next batch's loads are issued
// before the current batch is binned: the barrier does not drain the memory pipeline.
// (Measured and rejected: the image split over the shared memories of a 2-CTA cluster, 32-bit counters, both CTAs reading
// every ray: 5.25 ms for 1e8 rays; binning into the partner's half through distributed shared memory: 2.92 ms; the general
// kernel below: 2.27 ms.)
#include <atomic>
#include "optics_program.cuh"      // OpticsArgs, apply_program, bin_of (host + device)

namespace tt {

static constexpr int kSmemThreads = 512;
static constexpr int kSmemUnroll = 4;

// bin_of() on edges held in shared memory, the scale hoisted out of the ray loop: the two walks settle on the bin numpy's
// searchsorted finds wherever the first guess lands, so the result is bin_of()'s
__device__ __forceinline__ int bin_of_staged(double x, const double* e, int nb, double lo, double hi, double scale) {
    if (!(x >= lo && x <= hi)) return -1;
    int b = (int)((x - lo) * scale);
    b = b < 0 ? 0 : (b > nb - 1 ? nb - 1 : b);
    while (b > 0 && x < e[b]) --b;
    while (b < nb - 1 && x >= e[b + 1]) ++b;
    return b;
}

__global__ void __launch_bounds__(kSmemThreads, 1)
optics_hist_smem16_kernel(const double* __restrict__ rf_in, const double* __restrict__ xe, const double* __restrict__ ye,
                          unsigned long long* __restrict__ H, OpticsArgs A, int nwords, long niter) {
    extern __shared__ double smem_d[];
    double* sx = smem_d;                                     // nbx + 1 edges
    double* sy = sx + (A.nbx + 1);                           // nby + 1 edges
    unsigned* words = reinterpret_cast<unsigned*>(sy + (A.nby + 1));      // two 16-bit counters per word
    for (int i = threadIdx.x; i <= A.nbx; i += kSmemThreads) sx[i] = xe[i];
    for (int i = threadIdx.x; i <= A.nby; i += kSmemThreads) sy[i] = ye[i];
    for (int i = threadIdx.x; i < nwords; i += kSmemThreads) words[i] = 0u;
    __syncthreads();
    const double xlo = sx[0], xhi = sx[A.nbx], xs = (double)A.nbx / (xhi - xlo);
    const double ylo = sy[0], yhi = sy[A.nby], ys = (double)A.nby / (yhi - ylo);
    const long nthreads = (long)gridDim.x * kSmemThreads;
    const long t0 = (long)blockIdx.x * kSmemThreads + threadIdx.x;
    double cur[kSmemUnroll][4], nxt[kSmemUnroll][4];
#define TT_HIST_LOAD(buf, it)                                                                                    \
    _Pragma("unroll") for (int k = 0; k < kSmemUnroll; ++k) {                                                    \
        const long i = t0 + ((it) * kSmemUnroll + k) * nthreads;                                                 \
        if (i < A.np) {                                                                                          \
            buf[k][0] = __ldg(rf_in + i); buf[k][1] = __ldg(rf_in + A.np + i);                                   \
            buf[k][2] = __ldg(rf_in + 2 * A.np + i); buf[k][3] = __ldg(rf_in + 3 * A.np + i);                    \
        }                                                                                                        \
    }
    TT_HIST_LOAD(cur, 0L)
    for (long it = 0; it < niter; ++it) {                    // the same trip count for every thread: barrier inside
        if (it + 1 < niter) { TT_HIST_LOAD(nxt, it + 1) }
#pragma unroll
        for (int k = 0; k < kSmemUnroll; ++k) {
            const long i = t0 + (it * kSmemUnroll + k) * nthreads;
            if (i >= A.np) continue;
            double X = cur[k][0] * A.pos_scale, T = cur[k][1], Y = cur[k][2] * A.pos_scale, P = cur[k][3];
            apply_program(A, X, T, Y, P);
            const int ix = bin_of_staged(X, sx, A.nbx, xlo, xhi, xs), iy = bin_of_staged(Y, sy, A.nby, ylo, yhi, ys);
            if (ix >= 0 && iy >= 0) {
                const int b = iy * A.nbx + ix;
                const unsigned sh = (unsigned)(b & 1) * 16u;
                const unsigned old = atomicAdd(words + (b >> 1), 1u << sh);
                if (((old >> sh) & 0xFFFFu) == 0x7FFFu) {    // this add made it 0x8000: move them out
                    atomicSub(words + (b >> 1), 0x8000u << sh);
                    atomicAdd(&H[b], 0x8000ull);
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kSmemUnroll; ++k)
#pragma unroll
            for (int m = 0; m < 4; ++m) cur[k][m] = nxt[k][m];
    }
#undef TT_HIST_LOAD
    const int nbins = A.nbx * A.nby;
    for (int i = threadIdx.x; i < nwords; i += kSmemThreads) {
        const unsigned w = words[i];
        if (w & 0xFFFFu) atomicAdd(&H[2 * i], (unsigned long long)(w & 0xFFFFu));
        if ((w >> 16) && 2 * i + 1 < nbins) atomicAdd(&H[2 * i + 1], (unsigned long long)(w >> 16));
    }
}

// launches the privatised kernel if the image fits; returns TT_ERR_UNSUPPORTED (without setting an error) if it does not
static int launch_smem16_hist(const double* rf_in, const double* xe, const double* ye, unsigned long long* H, const OpticsArgs& A,
                              cudaStream_t s) {
    static const size_t kSmemMax = 227 * 1024;               // what a CTA may opt in to on sm_100
    const long nbins = (long)A.nbx * A.nby;
    const long nwords = (nbins + 1) / 2;
    const size_t smem = (size_t)(A.nbx + 1 + A.nby + 1) * sizeof(double) + (size_t)nwords * sizeof(unsigned);
    if (smem > kSmemMax) return TT_ERR_UNSUPPORTED;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    static std::atomic<int> attr_set[64];                    // per device: the attribute belongs to the device's copy of the kernel
    if (dev < 0 || dev >= 64) return TT_ERR_UNSUPPORTED;
    if (!attr_set[dev].load()) {
        if (cudaFuncSetAttribute(optics_hist_smem16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax) != cudaSuccess) {
            (void)cudaGetLastError();
            return TT_ERR_UNSUPPORTED;
        }
        attr_set[dev].store(1);
    }
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long per = (long)kSmemThreads * kSmemUnroll;
    long blocks = (A.np + per - 1) / per;
    if (blocks > sms) blocks = sms;                          // one CTA per SM (shared-memory limited)
    const long niter = (A.np + blocks * per - 1) / (blocks * per);
    optics_hist_smem16_kernel<<<(unsigned)blocks, kSmemThreads, smem, s>>>(rf_in, xe, ye, H, A, (int)nwords, niter);
    if (cudaPeekAtLastError() != cudaSuccess) {              // could not be launched: the general kernel does it
        (void)cudaGetLastError();
        return TT_ERR_UNSUPPORTED;
    }
    count_launch();
    return TT_OK;
}

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
    // the image alone, rays in storage order, and it fits a CTA's shared memory as 16-bit counters: privatised binning
    if (H_dev && !Hw_dev && !rf_out_dev && !perm_dev && np >= (1L << 16) &&
        launch_smem16_hist(rf_in_dev, xedges_dev, yedges_dev, H_dev, A, (cudaStream_t)stream) == TT_OK)
        return TT_OK;
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
