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
// value atomicAdd returns) makes its thread move 0x8000 counts to the global image, and a CTA-wide barrier per two batches of
// kSmemThreads x kSmemUnroll = 2560 rays bounds what the other threads can add in between (0x7FFF + 5120 < 0xFFFF; the
// subtraction lands before the barrier, so every counter is <= 0x7FFF at every barrier).  The next batch's loads are issued
// before the current batch is binned: the barrier does not drain the memory pipeline.
// (Measured and rejected: the image split over the shared memories of a 2-CTA cluster, 32-bit counters, both CTAs reading
// every ray: 5.25 ms for 1e8 rays; binning into the partner's half through distributed shared memory: 2.92 ms; the general
// kernel below: 2.27 ms.)
#include <atomic>
#include "optics_program.cuh"      // OpticsArgs, apply_program, bin_of (host + device)

namespace tt {

#ifndef TT_HIST_THREADS
#define TT_HIST_THREADS 640      // measured 384 / 512 / 640 / 768 threads: 1.22 / 1.02 / 0.91 / 1.03 ms per 1e8 rays (92 registers at 640; 768 spills)
#endif
static constexpr int kSmemThreads = TT_HIST_THREADS;
static constexpr int kSmemUnroll = 4;                        // rays per thread and batch
static_assert(2 * TT_HIST_THREADS * 4 < 0x8000, "the overflow protocol needs fewer than 0x8000 adds between two barriers");
static constexpr int kSmemChunk = kSmemThreads * kSmemUnroll;      // a batch = 2560 consecutive rays

// VEC: the four rows of rf_in are 16-byte aligned (even ray count): a thread loads its rays as two double2 per row
template <bool VEC>
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
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    // batch `it` of this CTA = rays [c * kSmemChunk, (c + 1) * kSmemChunk), c = it * gridDim.x + blockIdx.x; the thread's rays
    // in it: 2 t, 2 t + 1, kSmemChunk / 2 + 2 t, kSmemChunk / 2 + 2 t + 1 (constant offsets from one address per row)
    double bufa[4][kSmemUnroll], bufb[4][kSmemUnroll];       // [row][ray]
#define TT_HIST_LOAD(buf, it)                                                                                    \
    {                                                                                                            \
        const long first = ((it) * (long)gridDim.x + blockIdx.x) * kSmemChunk;                                   \
        const double* p = rf_in + first + 2 * threadIdx.x;                                                       \
        if (VEC && first + kSmemChunk <= A.np) {                                                                 \
            _Pragma("unroll") for (int m = 0; m < 4; ++m) {                                                      \
                const double2 v0 = __ldg(reinterpret_cast<const double2*>(p + m * A.np));                        \
                const double2 v1 = __ldg(reinterpret_cast<const double2*>(p + m * A.np + kSmemChunk / 2));       \
                buf[m][0] = v0.x; buf[m][1] = v0.y; buf[m][2] = v1.x; buf[m][3] = v1.y;                          \
            }                                                                                                    \
        } else {                                             /* the last batch, batches behind it, odd alignment */ \
            _Pragma("unroll") for (int k = 0; k < kSmemUnroll; ++k) {                                            \
                const long i = first + 2 * threadIdx.x + (k & 1) + (k >> 1) * (kSmemChunk / 2);                  \
                _Pragma("unroll") for (int m = 0; m < 4; ++m) buf[m][k] = i < A.np ? __ldg(rf_in + m * A.np + i) : nan;   \
            }                                                                                                    \
        }                                                                                                        \
    }
    // one batch: kSmemUnroll rays per thread side by side through the program (independent chains), then binned; a ray
    // behind the end is a NaN column and falls out at the bin search
#define TT_HIST_BIN(buf)                                                                                         \
    {                                                                                                            \
        _Pragma("unroll") for (int k = 0; k < kSmemUnroll; ++k) { buf[0][k] *= A.pos_scale; buf[2][k] *= A.pos_scale; }   \
        bool dead[kSmemUnroll];                                                                                  \
        run_program_n<kSmemUnroll>(A, buf[0], buf[1], buf[2], buf[3], dead);                                     \
        int b[kSmemUnroll];                                                                                      \
        _Pragma("unroll") for (int k = 0; k < kSmemUnroll; ++k) {                                                \
            /* a NaN in any row drops the ray (x, y: the bin search; theta, phi: here), unless the program is empty */ \
            if (A.n_ops > 0) dead[k] = dead[k] || buf[1][k] != buf[1][k] || buf[3][k] != buf[3][k];              \
            const int ix = bin_of_scaled<false>(buf[0][k], sx, A.nbx, xlo, xhi, xs);                             \
            const int iy = bin_of_scaled<false>(buf[2][k], sy, A.nby, ylo, yhi, ys);                             \
            b[k] = (!dead[k] && ix >= 0 && iy >= 0) ? iy * A.nbx + ix : -1;                                      \
        }                                                                                                        \
        _Pragma("unroll") for (int k = 0; k < kSmemUnroll; ++k) {                                                \
            if (b[k] < 0) continue;                                                                              \
            const unsigned sh = (unsigned)(b[k] & 1) * 16u;                                                      \
            const unsigned old = atomicAdd(words + (b[k] >> 1), 1u << sh);                                       \
            if (((old >> sh) & 0xFFFFu) == 0x7FFFu) { /* this add made it 0x8000: move them out */               \
                atomicSub(words + (b[k] >> 1), 0x8000u << sh);                                                   \
                atomicAdd(&H[b[k]], 0x8000ull);                                                                  \
            }                                                                                                    \
        }                                                                                                        \
    }
    // two batches per turn, the buffers taking turns (no register copies); every thread makes the same number of turns
    // (the barrier), batches behind the end of the rays are empty.  Adds between two barriers: 2 x kSmemChunk = 5120 < 0x8000.
    TT_HIST_LOAD(bufa, 0L)
    for (long it = 0; it < niter; it += 2) {
        TT_HIST_LOAD(bufb, it + 1)
        TT_HIST_BIN(bufa)
        TT_HIST_LOAD(bufa, it + 2)
        TT_HIST_BIN(bufb)
        __syncthreads();
    }
#undef TT_HIST_LOAD
#undef TT_HIST_BIN
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
        if (cudaFuncSetAttribute(optics_hist_smem16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax) != cudaSuccess ||
            cudaFuncSetAttribute(optics_hist_smem16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax) != cudaSuccess) {
            (void)cudaGetLastError();
            return TT_ERR_UNSUPPORTED;
        }
        attr_set[dev].store(1);
    }
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long chunks = (A.np + kSmemChunk - 1) / kSmemChunk;
    const long blocks = chunks < sms ? chunks : sms;         // one CTA per SM (shared-memory limited)
    const long niter = (chunks + blocks - 1) / blocks;
    const bool vec = (reinterpret_cast<uintptr_t>(rf_in) % 16 == 0) && (A.np % 2 == 0);
    if (vec) optics_hist_smem16_kernel<true><<<(unsigned)blocks, kSmemThreads, smem, s>>>(rf_in, xe, ye, H, A, (int)nwords, niter);
    else optics_hist_smem16_kernel<false><<<(unsigned)blocks, kSmemThreads, smem, s>>>(rf_in, xe, ye, H, A, (int)nwords, niter);
    if (cudaPeekAtLastError() != cudaSuccess) {              // could not be launched: the general kernel does it
        (void)cudaGetLastError();
        return TT_ERR_UNSUPPORTED;
    }
    count_launch();
    return TT_OK;
}

static constexpr int kThreads = 256;
static constexpr int kRaysPerThread = 8;
static constexpr int kGroup = 4;          // rays a thread takes through the program side by side
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
    double xlo = 0, xhi = 0, xs = 0, ylo = 0, yhi = 0, ys = 0;
    if (H || Hw) {
        xlo = ldg_f64(xe); xhi = ldg_f64(xe + A.nbx); xs = (double)A.nbx / (xhi - xlo);
        ylo = ldg_f64(ye); yhi = ldg_f64(ye + A.nby); ys = (double)A.nby / (yhi - ylo);
    }
    int bx[kRaysPerThread], by[kRaysPerThread];
    int mnx = 0x7fffffff, mny = 0x7fffffff;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
    for (int g = 0; g < kRaysPerThread; g += kGroup) {       // kGroup rays side by side: loads first, then the program
        long ray[kGroup];
        double x[kGroup], th[kGroup], y[kGroup], ph[kGroup];
#pragma unroll
        for (int k = 0; k < kGroup; ++k) {
            const long i = base + (long)(g + k) * kThreads + threadIdx.x;
            ray[k] = -1;
            x[k] = th[k] = y[k] = ph[k] = nan;
            if (i < A.np) {
                ray[k] = perm ? (long)perm[i] : i;
                x[k] = rf_in[ray[k]]; th[k] = rf_in[A.np + ray[k]]; y[k] = rf_in[2 * A.np + ray[k]]; ph[k] = rf_in[3 * A.np + ray[k]];
            }
        }
#pragma unroll
        for (int k = 0; k < kGroup; ++k) { x[k] *= A.pos_scale; y[k] *= A.pos_scale; }
        apply_program_n<kGroup>(A, x, th, y, ph);
#pragma unroll
        for (int k = 0; k < kGroup; ++k) {
            bx[g + k] = by[g + k] = -1;
            if (ray[k] < 0) continue;
            if (rf_out) {
                rf_out[ray[k]] = x[k]; rf_out[A.np + ray[k]] = th[k]; rf_out[2 * A.np + ray[k]] = y[k]; rf_out[3 * A.np + ray[k]] = ph[k];
            }
            if (H || Hw) {
                const int ix = bin_of_scaled<true>(x[k], xe, A.nbx, xlo, xhi, xs), iy = bin_of_scaled<true>(y[k], ye, A.nby, ylo, yhi, ys);
                if (ix >= 0 && iy >= 0) {
                    bx[g + k] = ix; by[g + k] = iy;
                    mnx = min(mnx, ix); mny = min(mny, iy);
                    // weighted image (numpy.histogram2d(weights=)): FP64 atomics straight to global
                    if (Hw) atomicAdd(&Hw[(size_t)iy * A.nbx + ix], weights[ray[k]]);
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
    prepare_program(A);                                      // -1/f, r^2: once, not per ray
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
