// Ray integrator and gradient look-up on RECTILINEAR grids: axes whose nodes are not equally spaced.
// The reference accepts any ascending coordinate arrays (RegularGridInterpolator + numpy.gradient with
// coordinates, particle_tracker.py:235-241); simulation cubes with stretched meshes take
// this path.  Same physics, outputs, status flags and edge-case rules as the uniform kernels of trace.cu:
//   * plane marching in the probing-axis coordinate w: every RK4 step starts and ends on a node plane (or one
//     of steps_per_cell sub-planes) of the w axis, so no step straddles the field's kink at a cell face in w;
//   * steep / backward / time-capped rays continue in an arc-length RK4 whose step is the smallest side of
//     the ray's current cell / steps_per_cell;
//   * positions are physical coordinates in FP64 (cells have different sizes, so the uniform kernels'
//     (cell, fraction) state does not carry over); the cell of a stage position is found by walking from
//     the ray's previous cell (rays move a fraction of a cell per step; the first look-up is a bisection).
// This is the general-geometry path, not the speed path: all arithmetic is FP64, the grid is read as
// float4 or double4.
#include "trace_axes_event.cuh"     // AxesArgs, find_cell, axes_event_ray (+ trace_common.cuh)

namespace tt {

// cell c in [0, n-2] with ax[c] <= x < ax[c+1] (clamped at both ends), walking from c
__device__ __forceinline__ int walk_cell(const double* __restrict__ ax, int n, double x, int c) {
    while (c > 0 && x < __ldg(ax + c)) --c;
    while (c < n - 2 && x >= __ldg(ax + c + 1)) ++c;
    return c;
}
struct RRay {
    double p[3];      // physical position (frame order)
    double d[3];      // v / c
    double s;         // path time c*t inside the cube
    int c[3];         // current cell per axis (look-up hint)
};

template <typename T>
struct Field {
    const typename GridT<T>::V4* grid;
    const AxesArgs& A;
    size_t plane;
    // gradient at a physical position (clamped into the cube); hint cells updated in place
    __device__ __forceinline__ void at(double pu, double pv, double pw, int c[3], double g[3]) const {
        double t[3];
        const double pp[3] = {pu, pv, pw};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            c[k] = walk_cell(A.ax[k], A.n[k], pp[k], c[k]);
            const double x0 = __ldg(A.ax[k] + c[k]), x1 = __ldg(A.ax[k] + c[k] + 1);
            double tt_ = (pp[k] - x0) / (x1 - x0);
            t[k] = tt_ < 0.0 ? 0.0 : (tt_ > 1.0 ? 1.0 : tt_);
        }
        const G3<double> r = lookup(c[0], c[1], c[2], t[0], t[1], t[2]);
        g[0] = r.x; g[1] = r.y; g[2] = r.z;
    }
    __device__ __forceinline__ G3<double> lookup(int cu, int cv, int cw, double tu, double tv, double tw) const {
        typedef typename GridT<T>::V4 V4;
        const V4* p = grid + ((size_t)cw * plane + (size_t)cv * A.n[0] + cu);
        const int nu = A.n[0];
        V4 c000 = GridT<T>::ld(p), c100 = GridT<T>::ld(p + 1), c010 = GridT<T>::ld(p + nu), c110 = GridT<T>::ld(p + nu + 1);
        p += plane;
        V4 c001 = GridT<T>::ld(p), c101 = GridT<T>::ld(p + 1), c011 = GridT<T>::ld(p + nu), c111 = GridT<T>::ld(p + nu + 1);
        G3<double> g;
#define TT_TRI(m)                                                                              \
    {                                                                                          \
        double a00 = fma(tu, (double)c100.m - (double)c000.m, (double)c000.m);                 \
        double a10 = fma(tu, (double)c110.m - (double)c010.m, (double)c010.m);                 \
        double a01 = fma(tu, (double)c101.m - (double)c001.m, (double)c001.m);                 \
        double a11 = fma(tu, (double)c111.m - (double)c011.m, (double)c011.m);                 \
        double b0 = fma(tv, a10 - a00, a00), b1 = fma(tv, a11 - a01, a01);                     \
        g.m = fma(tw, b1 - b0, b0);                                                            \
    }
        TT_TRI(x) TT_TRI(y) TT_TRI(z)
#undef TT_TRI
        return g;
    }
};

// RK4 step of size H in w (physical) from plane w0; false if a stage saw d_w <= 0
template <typename T>
__device__ __forceinline__ bool wstep(const Field<T>& F, RRay& r, double w0, double H) {
    const double half = 0.5 * H;
    double g[3];
    F.at(r.p[0], r.p[1], w0, r.c, g);
    bool ok = r.d[2] > 0.0;
    double q = 1.0 / r.d[2];
    const double aU = r.d[0] * q, aV = r.d[1] * q, a0 = g[0] * q, a1 = g[1] * q, a2 = g[2] * q, as = q;
    double du = fma(half, a0, r.d[0]), dv = fma(half, a1, r.d[1]), dw = fma(half, a2, r.d[2]);
    F.at(fma(half, aU, r.p[0]), fma(half, aV, r.p[1]), w0 + half, r.c, g);
    ok = ok && dw > 0.0; q = 1.0 / dw;
    const double bU = du * q, bV = dv * q, b0 = g[0] * q, b1 = g[1] * q, b2 = g[2] * q, bs = q;
    du = fma(half, b0, r.d[0]); dv = fma(half, b1, r.d[1]); dw = fma(half, b2, r.d[2]);
    F.at(fma(half, bU, r.p[0]), fma(half, bV, r.p[1]), w0 + half, r.c, g);
    ok = ok && dw > 0.0; q = 1.0 / dw;
    const double cU = du * q, cV = dv * q, c0 = g[0] * q, c1 = g[1] * q, c2 = g[2] * q, cs = q;
    du = fma(H, c0, r.d[0]); dv = fma(H, c1, r.d[1]); dw = fma(H, c2, r.d[2]);
    F.at(fma(H, cU, r.p[0]), fma(H, cV, r.p[1]), w0 + H, r.c, g);
    ok = ok && dw > 0.0; q = 1.0 / dw;
    const double eU = du * q, eV = dv * q, e0 = g[0] * q, e1 = g[1] * q, e2 = g[2] * q, es = q;
    const double h6 = H * (1.0 / 6.0);
    r.p[0] = fma(h6, aU + 2.0 * (bU + cU) + eU, r.p[0]);
    r.p[1] = fma(h6, aV + 2.0 * (bV + cV) + eV, r.p[1]);
    r.d[0] = fma(h6, a0 + 2.0 * (b0 + c0) + e0, r.d[0]);
    r.d[1] = fma(h6, a1 + 2.0 * (b1 + c1) + e1, r.d[1]);
    r.d[2] = fma(h6, a2 + 2.0 * (b2 + c2) + e2, r.d[2]);
    r.s = fma(h6, as + 2.0 * (bs + cs) + es, r.s);
    r.p[2] = w0 + H;
    return ok;
}

// RK4 step of length ds in path time
template <typename T>
__device__ __forceinline__ void pstep(const Field<T>& F, RRay& r, double ds) {
    const double half = 0.5 * ds;
    double g1[3], g2[3], g3[3], g4[3], d2[3], d3[3], d4[3];
    F.at(r.p[0], r.p[1], r.p[2], r.c, g1);
#pragma unroll
    for (int k = 0; k < 3; ++k) d2[k] = fma(half, g1[k], r.d[k]);
    F.at(fma(half, r.d[0], r.p[0]), fma(half, r.d[1], r.p[1]), fma(half, r.d[2], r.p[2]), r.c, g2);
#pragma unroll
    for (int k = 0; k < 3; ++k) d3[k] = fma(half, g2[k], r.d[k]);
    F.at(fma(half, d2[0], r.p[0]), fma(half, d2[1], r.p[1]), fma(half, d2[2], r.p[2]), r.c, g3);
#pragma unroll
    for (int k = 0; k < 3; ++k) d4[k] = fma(ds, g3[k], r.d[k]);
    F.at(fma(ds, d3[0], r.p[0]), fma(ds, d3[1], r.p[1]), fma(ds, d3[2], r.p[2]), r.c, g4);
    const double s6 = ds * (1.0 / 6.0);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        r.p[k] = fma(s6, r.d[k] + 2.0 * (d2[k] + d3[k]) + d4[k], r.p[k]);
        r.d[k] = fma(s6, g1[k] + 2.0 * (g2[k] + g3[k]) + g4[k], r.d[k]);
    }
    r.s += ds;
}

// fraction of the chord p_old -> p_new at which the coordinate leaves [lo, hi]; 2 if it does not
__device__ __forceinline__ double leave_frac(double p_old, double p_new, double lo, double hi) {
    if (p_new < lo) return (lo - p_old) / (p_new - p_old);
    if (p_new > hi) return (hi - p_old) / (p_new - p_old);
    return 2.0;
}

template <typename T>
__global__ void __launch_bounds__(128) trace_axes_kernel(const typename GridT<T>::V4* __restrict__ grid,
                                                         const double* __restrict__ s0,
                                                         const uint32_t* __restrict__ perm, double* __restrict__ rf,
                                                         double* __restrict__ sf,
                                                         unsigned long long* __restrict__ ray_steps,
                                                         uint8_t* __restrict__ status, AxesArgs A,
                                                         const unsigned int* __restrict__ any_deferred = nullptr) {
    // second pass behind trace_axes_event_kernel (any_deferred != null): only the rays it handed over
    if (any_deferred && *any_deferred == 0u) return;
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned steps = 0;
    bool mine = tid < A.np;
    if (mine && any_deferred) mine = status[perm ? (long)perm[tid] : tid] == TT_RAY_DEFERRED;
    if (mine) {
        const long ray = perm ? (long)perm[tid] : tid;
        double P[3], D[3], lo[3], hi[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            P[k] = s0[(size_t)A.fa[k] * A.np + ray];
            D[k] = s0[(size_t)(3 + A.fa[k]) * A.np + ray] * (1.0 / kC);
            lo[k] = __ldg(A.ax[k]); hi[k] = __ldg(A.ax[k] + A.n[k] - 1);
        }
        int st = 0;
        double s_acc = 0.0;
        RRay r;
#pragma unroll
        for (int k = 0; k < 3; ++k) { r.p[k] = P[k]; r.d[k] = D[k]; }
        // ---- prologue: free flight to the cube if launched outside ----------------------------------
        // (a NaN / inf launch state is "a ray that misses": see trace.cu)
        bool finite = true;
#pragma unroll
        for (int k = 0; k < 3; ++k) finite = finite && isfinite(P[k]) && isfinite(D[k]);
        bool inside = finite;
#pragma unroll
        for (int k = 0; k < 3; ++k) inside = inside && P[k] >= lo[k] && P[k] <= hi[k];
        if (!inside) {
            double t_in = 0.0, t_out = 1e300;
            bool hit = finite;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (D[k] == 0.0) {
                    hit = hit && P[k] >= lo[k] && P[k] <= hi[k];
                } else {
                    const double ta = (lo[k] - P[k]) / D[k], tb = (hi[k] - P[k]) / D[k];
                    t_in = fmax(t_in, fmin(ta, tb));
                    t_out = fmin(t_out, fmax(ta, tb));
                }
            }
            hit = hit && t_in <= t_out && t_in < A.s_max;
            if (hit) {
#pragma unroll
                for (int k = 0; k < 3; ++k) r.p[k] = fmin(fmax(fma(D[k], t_in, P[k]), lo[k]), hi[k]);
                s_acc = t_in;
            } else {
                st = TT_RAY_MISSED;
            }
        }
        r.s = 0.0;
        const double s_left0 = A.s_max - s_acc;
        Field<T> F{grid, A, (size_t)A.n[0] * A.n[1]};
        const double* axw = A.ax[2];
        const int nw = A.n[2];
        bool general = false;
        if (!(st & TT_RAY_MISSED)) {
#pragma unroll
            for (int k = 0; k < 3; ++k) r.c[k] = find_cell(A.ax[k], A.n[k], r.p[k]);
            if (!(r.d[2] > TT_MARCH_MIN_DW)) general = true;
            if (!general && (hi[2] - r.p[2]) > TT_MARCH_MIN_DW * s_left0) general = true;
            bool alive = true;
            // ---- plane marching ---------------------------------------------------------------------
            if (!general) {
                int k = r.c[2];
                if (r.p[2] >= hi[2]) k = nw - 1;                     // already on the far face
                while (k < nw - 1) {
                    const double w_lo = __ldg(axw + k), w_hi = __ldg(axw + k + 1), hk = w_hi - w_lo;
                    int j = (int)((r.p[2] - w_lo) / hk * (double)A.spc);      // sub-plane interval containing w
                    j = j < 0 ? 0 : (j > A.spc - 1 ? A.spc - 1 : j);
                    bool stop = false;
                    for (; j < A.spc; ++j) {
                        const double wa = r.p[2];
                        const double wb = (j + 1 == A.spc) ? w_hi : fma((double)(j + 1) / (double)A.spc, hk, w_lo);
                        const double H = wb - wa;
                        const RRay old = r;
                        const bool ok = wstep<T>(F, r, wa, H);
                        r.p[2] = wb;
                        ++steps;
                        if (!ok || !(r.d[2] > TT_MARCH_MIN_DW) || !(r.s <= s_left0)) {
                            r = old; --steps;
                            general = true; stop = true;
                            break;
                        }
                        double lam = fmin(leave_frac(old.p[0], r.p[0], lo[0], hi[0]), leave_frac(old.p[1], r.p[1], lo[1], hi[1]));
                        if (lam <= 1.0) {      // side exit: re-step to the face, freeze there
                            r = old;
                            lam = lam < 0.0 ? 0.0 : lam;
                            wstep<T>(F, r, wa, lam * H);
                            r.p[0] = fmin(fmax(r.p[0], lo[0]), hi[0]);
                            r.p[1] = fmin(fmax(r.p[1], lo[1]), hi[1]);
                            st |= TT_RAY_EXIT_SIDE;
                            alive = false; stop = true;
                            break;
                        }
                    }
                    if (stop) break;
                    ++k;
                    r.c[2] = k < nw - 1 ? k : nw - 2;
                }
                if (alive && !general) { r.p[2] = hi[2]; st |= TT_RAY_EXIT_FACE; alive = false; }
            }
            // ---- general arc-length integrator ------------------------------------------------------
            if (alive && general) {
                st |= TT_RAY_GENERAL;
                for (long it = 0; it < (1L << 40); ++it) {
                    const double left = s_left0 - r.s;
                    if (!(left > 0.0)) { st |= TT_RAY_TIME_CAP; break; }
                    double hmin = 1e300;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        r.c[k] = walk_cell(A.ax[k], A.n[k], r.p[k], r.c[k]);
                        hmin = fmin(hmin, __ldg(A.ax[k] + r.c[k] + 1) - __ldg(A.ax[k] + r.c[k]));
                    }
                    const double ds0 = hmin / (double)A.spc;
                    const double ds = ds0 < left ? ds0 : left;
                    const RRay old = r;
                    pstep<T>(F, r, ds);
                    ++steps;
                    const double lu = leave_frac(old.p[0], r.p[0], lo[0], hi[0]);
                    const double lv = leave_frac(old.p[1], r.p[1], lo[1], hi[1]);
                    const double lw = leave_frac(old.p[2], r.p[2], lo[2], hi[2]);
                    double lam = fmin(lu, fmin(lv, lw));
                    if (lam <= 1.0) {
                        const bool far_face = (lw <= lu && lw <= lv) && r.p[2] > old.p[2];
                        r = old;
                        lam = lam < 0.0 ? 0.0 : lam;
                        if (lam > 0.0) pstep<T>(F, r, lam * ds);
#pragma unroll
                        for (int k = 0; k < 3; ++k) r.p[k] = fmin(fmax(r.p[k], lo[k]), hi[k]);
                        st |= far_face ? TT_RAY_EXIT_FACE : TT_RAY_EXIT_SIDE;
                        break;
                    }
                    if (ds < ds0) { st |= TT_RAY_TIME_CAP; break; }
                }
            }
        }
        // ---- epilogue: ray_at_exit (particle_tracker.py:345-380) and the state at time T -------------
        double Pf[3], Vf[3];
        if (st & TT_RAY_MISSED) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { Pf[k] = P[k]; Vf[k] = D[k] * kC; }
            s_acc = 0.0;
        } else {
#pragma unroll
            for (int k = 0; k < 3; ++k) { Pf[k] = r.p[k]; Vf[k] = r.d[k] * kC; }
            s_acc += r.s;
        }
        const double tb = (Pf[2] - A.extent) / Vf[2];
        rf[0 * A.np + ray] = Pf[0] - Vf[0] * tb;
        rf[1 * A.np + ray] = atan(Vf[0] / Vf[2]);
        rf[2 * A.np + ray] = Pf[1] - Vf[1] * tb;
        rf[3 * A.np + ray] = atan(Vf[1] / Vf[2]);
        if (sf) {
            const double t_rest = (A.s_max - s_acc) / kC;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                sf[(size_t)A.fa[k] * A.np + ray] = Pf[k] + Vf[k] * t_rest;
                sf[(size_t)(3 + A.fa[k]) * A.np + ray] = Vf[k];
            }
        }
        if (status) status[ray] = (uint8_t)st;
    }
    if (ray_steps) {
        unsigned v = steps;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(ray_steps, (unsigned long long)v);
    }
}

// event marching on a rectilinear grid: one ray per thread, body in trace_axes_event.cuh (host + device)
template <typename T>
__global__ void __launch_bounds__(128, sizeof(T) == 8 ? 3 : 4)
trace_axes_event_kernel(const typename GridT<T>::V4* __restrict__ grid, const double* __restrict__ s0,
                        const uint32_t* __restrict__ perm, double* __restrict__ rf, double* __restrict__ sf,
                        unsigned long long* __restrict__ ray_steps, uint8_t* __restrict__ status, AxesArgs A,
                        unsigned int* __restrict__ any_deferred) {
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned steps = 0;
    if (tid < A.np) {
        const long ray = perm ? (long)perm[tid] : tid;
        bool deferred = false;
        steps = axes_event_ray<T>(grid, s0, ray, rf, sf, status, A, deferred);
        if (deferred) *any_deferred = 1u;           // (benign race: everybody stores 1)
    }
    if (ray_steps) {
        unsigned v = steps;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(ray_steps, (unsigned long long)v);
    }
}

// ElectronCube.dndr on a rectilinear grid: zero outside, faces inclusive (scipy _rgi.py:635-642)
template <typename T>
__global__ void dndr_axes_kernel(const typename GridT<T>::V4* __restrict__ grid, AxesArgs A,
                                 const double* __restrict__ pos, long npts, double* __restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npts) return;
    double g3[3] = {0.0, 0.0, 0.0};
    double p[3];
    bool inside = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        p[k] = pos[(size_t)A.fa[k] * npts + i];
        inside = inside && !(p[k] < __ldg(A.ax[k])) && !(p[k] > __ldg(A.ax[k] + A.n[k] - 1)) && p[k] == p[k];
    }
    if (inside) {
        int c[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) c[k] = find_cell(A.ax[k], A.n[k], p[k]);
        Field<T> F{grid, A, (size_t)A.n[0] * A.n[1]};
        F.at(p[0], p[1], p[2], c, g3);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) out[(size_t)A.fa[k] * npts + i] = g3[k] * (kC * kC);
}

static int fill_axes(AxesArgs& A, const int n_xyz[3], const double* x_dev, const double* y_dev, const double* z_dev,
                     int par) {
    TT_REQUIRE(n_xyz && x_dev && y_dev && z_dev, "null geometry pointer");
    TT_REQUIRE(par >= 0 && par <= 2, "par must be 0, 1 or 2 (got %d)", par);
    const double* axes[3] = {x_dev, y_dev, z_dev};
    Frame f = frame_of(par);
    for (int k = 0; k < 3; ++k) {
        A.fa[k] = f.a[k];
        A.n[k] = n_xyz[f.a[k]];
        A.ax[k] = axes[f.a[k]];
        TT_REQUIRE(A.n[k] >= 2, "every axis needs >= 2 points");
    }
    return TT_OK;
}

}  // namespace tt

extern "C" int tt_trace_axes(const tt_trace_params* p, const double* x_dev, const double* y_dev, const double* z_dev,
                             const void* grid4_dev, const double* s0_dev, long np, const uint32_t* perm_dev,
                             double* rf_dev, double* sf_dev, unsigned long long* ray_steps_dev, uint8_t* status_dev,
                             tt_stream_t stream) {
    using namespace tt;
    TT_REQUIRE(p && grid4_dev && s0_dev && rf_dev, "tt_trace_axes: null pointer");
    TT_REQUIRE(np >= 0, "tt_trace_axes: negative ray count");
    TT_REQUIRE(p->dtype == TT_F32 || p->dtype == TT_F64, "tt_trace_axes: dtype must be TT_F32 or TT_F64");
    TT_REQUIRE(p->steps_per_cell >= 1 && p->steps_per_cell <= 1024, "tt_trace_axes: steps_per_cell out of range");
    TT_REQUIRE(p->s_max > 0 && p->extent == p->extent, "tt_trace_axes: s_max must be > 0");
    TT_REQUIRE(np < (1L << 32) || !perm_dev, "tt_trace_axes: perm is 32-bit; trace in bundles of < 2^32 rays");
    AxesArgs A;
    int rc = fill_axes(A, p->n_xyz, x_dev, y_dev, z_dev, p->par);
    if (rc) return rc;
    A.extent = p->extent; A.s_max = p->s_max; A.spc = p->steps_per_cell; A.np = np;
    if (np == 0) return TT_OK;
    const int block = 128;
    const long blocks = (np + block - 1) / block;
    TT_REQUIRE(blocks < (1L << 31), "tt_trace_axes: too many rays for one launch");
    cudaStream_t s = (cudaStream_t)stream;
    // variant: 0 = auto (event marching when a status buffer is given), 3 = event marching (needs status_dev),
    //          1 / 2 = the gather kernel alone
    TT_REQUIRE(p->variant >= 0 && p->variant <= 3, "tt_trace_axes: unknown kernel variant %d", p->variant);
    TT_REQUIRE(p->variant != 3 || status_dev, "tt_trace_axes: event marching (variant 3) needs status_dev");
    unsigned int* flag = nullptr;
    if ((p->variant == 0 || p->variant == 3) && status_dev) {
        // stream-ordered 4-byte scratch flag: "did the event kernel defer any ray?"
        flag = scratch_flag(s);
        if (flag) {         // (no scratch: the gather kernel does everything)
            if (p->dtype == TT_F32)
                trace_axes_event_kernel<float><<<(unsigned)blocks, block, 0, s>>>((const float4*)grid4_dev, s0_dev, perm_dev, rf_dev,
                                                                                  sf_dev, ray_steps_dev, status_dev, A, flag);
            else
                trace_axes_event_kernel<double><<<(unsigned)blocks, block, 0, s>>>((const double4*)grid4_dev, s0_dev, perm_dev, rf_dev,
                                                                                   sf_dev, ray_steps_dev, status_dev, A, flag);
            int rc2 = launch_check("trace_axes_event_kernel");
            if (rc2) { cudaFreeAsync(flag, s); return rc2; }
        }
    }
    if (p->dtype == TT_F32)
        trace_axes_kernel<float><<<(unsigned)blocks, block, 0, s>>>((const float4*)grid4_dev, s0_dev, perm_dev, rf_dev, sf_dev,
                                                                    ray_steps_dev, status_dev, A, flag);
    else
        trace_axes_kernel<double><<<(unsigned)blocks, block, 0, s>>>((const double4*)grid4_dev, s0_dev, perm_dev, rf_dev, sf_dev,
                                                                     ray_steps_dev, status_dev, A, flag);
    if (flag) cudaFreeAsync(flag, s);
    return launch_check("trace_axes_kernel");
}

extern "C" int tt_dndr_axes(const void* grid4_dev, int grid_dtype, const int n_xyz[3], const double* x_dev,
                            const double* y_dev, const double* z_dev, int par, const double* pos_dev, long npts,
                            double* out_dev, tt_stream_t stream) {
    using namespace tt;
    TT_REQUIRE(grid4_dev && pos_dev && out_dev, "tt_dndr_axes: null pointer");
    TT_REQUIRE(grid_dtype == TT_F32 || grid_dtype == TT_F64, "tt_dndr_axes: dtype must be TT_F32 or TT_F64");
    AxesArgs A;
    int rc = fill_axes(A, n_xyz, x_dev, y_dev, z_dev, par);
    if (rc) return rc;
    A.extent = 0; A.s_max = 0; A.spc = 1; A.np = npts;
    if (npts <= 0) return TT_OK;
    const int block = 256;
    const long blocks = (npts + block - 1) / block;
    cudaStream_t s = (cudaStream_t)stream;
    if (grid_dtype == TT_F32)
        dndr_axes_kernel<float><<<(unsigned)blocks, block, 0, s>>>((const float4*)grid4_dev, A, pos_dev, npts, out_dev);
    else
        dndr_axes_kernel<double><<<(unsigned)blocks, block, 0, s>>>((const double4*)grid4_dev, A, pos_dev, npts, out_dev);
    return launch_check("dndr_axes_kernel");
}
