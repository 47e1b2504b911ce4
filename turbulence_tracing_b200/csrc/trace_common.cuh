// Device helpers shared by the trace kernels (trace.cu: gather / cell-cache kernels and the C ABI;
// trace_event.cu: event-marching kernels).  See trace.cu for the formulation.
#pragma once
#include "common.cuh"

namespace tt {

static constexpr double kC = 299792458.0;      // scipy.constants.c (particle_tracker.py:119)
#define TT_MARCH_MIN_DW 0.75
#define TT_RAY_DEFERRED 0xFF       // internal status: ray left to the second-pass kernel
#ifndef TT_TRACE_MIN_BLOCKS
#define TT_TRACE_MIN_BLOCKS 3      // CTAs of 128 threads per SM the register allocation aims for
#endif

struct TraceArgs {
    int n[3];        // nu, nv, nw
    double o[3];     // origin per frame axis
    double h[3];     // spacing per frame axis
    int fa[3];       // frame axis -> xyz row
    double extent, s_max;
    int spc;
    long np;
    // precomputed for the event kernels (operands straight from the constant bank: no in-loop 64-bit
    // multiplies, no double->float conversions)
    long long plane_elems;   // nu * nv
    float hwf, ruf, rvf;     // (float) h_w, h_w/h_u, h_w/h_v
    unsigned int list_cap;        // second pass: capacity of the compacted list of deferred rays (0: no list); the flag-scan
                                  // kernel then only runs when more rays than that were deferred (any_deferred[1] > list_cap)
    unsigned int* any_deferred;   // device flag (nullable): set by an event kernel that deferred a ray, so
                                  // that the second pass can return at once when there is nothing to do
};

template <typename T> struct GridT;
template <> struct GridT<float> {
    typedef float4 V4;
    static TT_HD float4 ld(const float4* p) {
#ifdef __CUDA_ARCH__
        return __ldg(p);
#else
        return *p;
#endif
    }
    // the three gradient lanes only (8 + 4 bytes).  The event kernels prefetch the next plane's corners a whole
    // iteration ahead: with a 16-byte load the unused 4th lane is a DEAD destination register that ptxas hands
    // out as a temporary, and the first write to it has to wait for the load in flight (write-after-write) --
    // measured as 11.6 % of the kernel's stall samples (profiles/r02_trace_v5_c3_*) on an FFMA that does not
    // even consume the loaded data
    static TT_HD float4 ld3(const float4* p) {
#ifdef __CUDA_ARCH__
        const float2 a = __ldg(reinterpret_cast<const float2*>(p));
        const float b = __ldg(reinterpret_cast<const float*>(p) + 2);
        return make_float4(a.x, a.y, b, 0.f);
#else
        return make_float4(p->x, p->y, p->z, 0.f);
#endif
    }
};
template <> struct GridT<double> {
    typedef double4 V4;
    static TT_HD double4 ld(const double4* p) {
#ifdef __CUDA_ARCH__
        const double2* q = reinterpret_cast<const double2*>(p);
        double2 a = __ldg(q), b = __ldg(q + 1);
        return make_double4(a.x, a.y, b.x, b.y);
#else
        return *p;
#endif
    }
    static TT_HD double4 ld3(const double4* p) {
#ifdef __CUDA_ARCH__
        const double2 a = __ldg(reinterpret_cast<const double2*>(p));
        const double b = __ldg(reinterpret_cast<const double*>(p) + 2);
        return make_double4(a.x, a.y, b, 0.0);
#else
        return make_double4(p->x, p->y, p->z, 0.0);
#endif
    }
};

template <typename T> TT_HD T tfloor(T x);
template <> TT_HD float tfloor<float>(float x) { return floorf(x); }
template <> TT_HD double tfloor<double>(double x) { return floor(x); }
template <typename T> TT_HD T tfma(T a, T b, T c);
template <> TT_HD float tfma<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <> TT_HD double tfma<double>(double a, double b, double c) { return fma(a, b, c); }

// cell/fraction of coordinate (i + f) clamped into [0, n-1]; the upper face is cell n-2, t = 1
// (as scipy's find_indices does for x == grid[-1]).
template <typename T>
TT_HD void cell_of(int i, T f, int n, int& c, T& t) {
    T fl = tfloor(f);
    c = i + (int)fl;
    t = f - fl;
    if (c < 0) { c = 0; t = T(0); }
    if (c > n - 2) { c = n - 2; t = T(1); }
}

template <typename T> struct G3 { T x, y, z; };

// trilinear gradient in cell (cu, cv, cw) at fractions (tu, tv, tw)
template <typename T>
TT_HD G3<T> trilinear(const typename GridT<T>::V4* __restrict__ grid, int nu,
                                           size_t plane, int cu, int cv, int cw, T tu, T tv, T tw) {
    typedef typename GridT<T>::V4 V4;
    const V4* p = grid + ((size_t)cw * plane + (size_t)cv * nu + cu);
    V4 c000 = GridT<T>::ld(p), c100 = GridT<T>::ld(p + 1);
    V4 c010 = GridT<T>::ld(p + nu), c110 = GridT<T>::ld(p + nu + 1);
    p += plane;
    V4 c001 = GridT<T>::ld(p), c101 = GridT<T>::ld(p + 1);
    V4 c011 = GridT<T>::ld(p + nu), c111 = GridT<T>::ld(p + nu + 1);
    G3<T> g;
#define TT_TRI(m)                                                        \
    {                                                                    \
        T a00 = tfma(tu, c100.m - c000.m, c000.m);                       \
        T a10 = tfma(tu, c110.m - c010.m, c010.m);                       \
        T a01 = tfma(tu, c101.m - c001.m, c001.m);                       \
        T a11 = tfma(tu, c111.m - c011.m, c011.m);                       \
        T b0 = tfma(tv, a10 - a00, a00);                                 \
        T b1 = tfma(tv, a11 - a01, a01);                                 \
        g.m = tfma(tw, b1 - b0, b0);                                     \
    }
    TT_TRI(x) TT_TRI(y) TT_TRI(z)
#undef TT_TRI
    return g;
}

// ---- the field of ONE cell as a trilinear polynomial per component (event-marching kernels) ----------
template <typename T>
struct Tri {           // trilinear polynomial of one component in one cell
    T a, b, c, d, a1, b1, c1, d1;
};
// evaluation is "w first": the bilinear coefficients at w-fraction fw (4 FMA), then the bilinear form
// (3 FMA).  RK4 stages 2 and 3 share their w-fraction, so their coefficients are formed once.
template <typename T>
struct Bil {
    T a, b, c, d;
};
template <typename T>
TT_HD Bil<T> tri_at(const Tri<T>& q, T fw) {
    Bil<T> r;
    r.a = tfma(fw, q.a1, q.a); r.b = tfma(fw, q.b1, q.b); r.c = tfma(fw, q.c1, q.c); r.d = tfma(fw, q.d1, q.d);
    return r;
}
template <typename T>
TT_HD T bil_eval(const Bil<T>& q, T tu, T tv) {
    return tfma(tv, tfma(tu, q.d, q.c), tfma(tu, q.b, q.a));
}
// coefficients of plane 0 from its 4 corners; primed = plane 1 minus plane 0
template <typename T>
TT_HD void tri_set(Tri<T>& q, T c00, T c10, T c01, T c11, T e00, T e10, T e01, T e11) {
    q.a = c00; q.b = c10 - c00; q.c = c01 - c00; q.d = (c11 - c01) - q.b;
    T ea = e00, eb = e10 - e00, ec = e01 - e00, ed = (e11 - e01) - eb;
    q.a1 = ea - q.a; q.b1 = eb - q.b; q.c1 = ec - q.c; q.d1 = ed - q.d;
}
// advance one plane: plane 1 becomes plane 0, (n00..n11) are the corners of the new plane 1
template <typename T>
TT_HD void tri_advance(Tri<T>& q, T n00, T n10, T n01, T n11) {
    q.a += q.a1; q.b += q.b1; q.c += q.c1; q.d += q.d1;
    T eb = n10 - n00;
    q.a1 = n00 - q.a; q.b1 = eb - q.b; q.c1 = (n01 - n00) - q.c; q.d1 = ((n11 - n01) - eb) - q.d;
}

// the two halves of tri_set / tri_advance (same operations on the same operands): base plane from its 4 corners or by
// stepping one plane, then the primed half from the 4 corners of the far plane
template <typename T>
TT_HD void tri_base(Tri<T>& q, T c00, T c10, T c01, T c11) {
    q.a = c00; q.b = c10 - c00; q.c = c01 - c00; q.d = (c11 - c01) - q.b;
}
template <typename T>
TT_HD void tri_base_step(Tri<T>& q) {
    q.a += q.a1; q.b += q.b1; q.c += q.c1; q.d += q.d1;
}
template <typename T>
TT_HD void tri_primed(Tri<T>& q, T n00, T n10, T n01, T n11) {
    T eb = n10 - n00;
    q.a1 = n00 - q.a; q.b1 = eb - q.b; q.c1 = (n01 - n00) - q.c; q.d1 = ((n11 - n01) - eb) - q.d;
}

template <typename T>
struct Ray {
    int iu, iv, iw;
    T fu, fv, fw;     // fractions relative to (iu, iv, iw); may be un-normalised after a step
    T du, dv, dw;
    T s;              // path time c*t accumulated inside the cube
};

template <typename T>
struct Consts {
    int nu, nv, nw;
    size_t plane;
    T ru, rv, hw;     // hw/hu, hw/hv, hw  (plane marching)
    T iu_, iv_, iw_;  // 1/hu, 1/hv, 1/hw  (arc-length stepping)
};

// ---- passive quantities carried along the ray (BASELINE config 4; no implementation in the reference
// checkout -- only call sites, example_kitchensink.py:72-101 -- so parity is unpinned; textbook forms):
//   phase          dphi/ds   = (omega/c) (n - 1),        n = sqrt(1 - ne/nc)
//   Faraday        dalpha/ds = V ne (B . d),              V = e^3 lambda^2 / (8 pi^2 eps0 me^2 c^3)
//   inv. brems.    dln(a)/ds = -kappa / 2                 (kappa: energy absorption coefficient, 1/m)
// with s = c t and d = v/c.  They do not act back on the trajectory, so the RK4 stages of the ray are
// reused as a Simpson quadrature (weights 1, 2, 2, 1).  aux4 = (B_u, B_v, B_w, kappa) on the same grid
// layout as the gradient grid, whose 4th lane is ne/nc.  Accumulated in FP64; scaled by the constants
// in the epilogue.
template <typename T>
struct AuxCtx {
    const typename GridT<T>::V4* aux4;     // may be null: phase only
    double phase, farad, absorb;           // integrals of (n-1), (ne/nc)(B.d), kappa over s
};

template <typename T>
TT_HD void trilinear_w(const typename GridT<T>::V4* __restrict__ grid, int nu, size_t plane,
                                            int cu, int cv, int cw, T tu, T tv, T tw, T& w_only) {
    typedef typename GridT<T>::V4 V4;
    const V4* p = grid + ((size_t)cw * plane + (size_t)cv * nu + cu);
    T c000 = GridT<T>::ld(p).w, c100 = GridT<T>::ld(p + 1).w, c010 = GridT<T>::ld(p + nu).w, c110 = GridT<T>::ld(p + nu + 1).w;
    p += plane;
    T c001 = GridT<T>::ld(p).w, c101 = GridT<T>::ld(p + 1).w, c011 = GridT<T>::ld(p + nu).w, c111 = GridT<T>::ld(p + nu + 1).w;
    T a00 = tfma(tu, c100 - c000, c000), a10 = tfma(tu, c110 - c010, c010);
    T a01 = tfma(tu, c101 - c001, c001), a11 = tfma(tu, c111 - c011, c011);
    T b0 = tfma(tv, a10 - a00, a00), b1 = tfma(tv, a11 - a01, a01);
    w_only = tfma(tw, b1 - b0, b0);
}

template <typename T>
TT_HD void trilinear4(const typename GridT<T>::V4* __restrict__ grid, int nu, size_t plane,
                                           int cu, int cv, int cw, T tu, T tv, T tw, T& x, T& y, T& z, T& w) {
    typedef typename GridT<T>::V4 V4;
    const V4* p = grid + ((size_t)cw * plane + (size_t)cv * nu + cu);
    V4 c000 = GridT<T>::ld(p), c100 = GridT<T>::ld(p + 1), c010 = GridT<T>::ld(p + nu), c110 = GridT<T>::ld(p + nu + 1);
    p += plane;
    V4 c001 = GridT<T>::ld(p), c101 = GridT<T>::ld(p + 1), c011 = GridT<T>::ld(p + nu), c111 = GridT<T>::ld(p + nu + 1);
#define TT_TRI4(m, out)                                                  \
    {                                                                    \
        T a00 = tfma(tu, c100.m - c000.m, c000.m);                       \
        T a10 = tfma(tu, c110.m - c010.m, c010.m);                       \
        T a01 = tfma(tu, c101.m - c001.m, c001.m);                       \
        T a11 = tfma(tu, c111.m - c011.m, c011.m);                       \
        T b0 = tfma(tv, a10 - a00, a00);                                 \
        T b1 = tfma(tv, a11 - a01, a01);                                 \
        out = tfma(tw, b1 - b0, b0);                                     \
    }
    TT_TRI4(x, x) TT_TRI4(y, y) TT_TRI4(z, z) TT_TRI4(w, w)
#undef TT_TRI4
}

// integrands at one stage, multiplied by `scale` (ds/dW = hw/dw when marching in W, 1 in path time)
template <typename T>
TT_HD void aux_rates(const AuxCtx<T>& ctx, const typename GridT<T>::V4* __restrict__ grid,
                                          const Consts<T>& C, int cu, int cv, int cw, T tu, T tv, T tw, T du, T dv,
                                          T dw, T scale, double& fp, double& ff, double& fa) {
    T nn;
    trilinear_w<T>(grid, C.nu, C.plane, cu, cv, cw, tu, tv, tw, nn);       // ne/nc
    const double nr = 1.0 - (double)nn;
    fp = (sqrt(nr > 0.0 ? nr : 0.0) - 1.0) * (double)scale;
    ff = 0.0; fa = 0.0;
    if (ctx.aux4) {
        T bu, bv, bw, kap;
        trilinear4<T>(ctx.aux4, C.nu, C.plane, cu, cv, cw, tu, tv, tw, bu, bv, bw, kap);
        ff = (double)nn * ((double)bu * du + (double)bv * dv + (double)bw * dw) * (double)scale;
        fa = (double)kap * (double)scale;
    }
}

// One RK4 step in W from fraction fwa to fwa + h inside w-cell k (0 <= fwa, fwa + h <= 1).
// Updates fu, fv (un-normalised), d and s of r.  Returns false if a stage saw d_w <= 0.
template <typename T, bool AUX = false>
TT_HD bool zstep(const typename GridT<T>::V4* __restrict__ grid, const Consts<T>& C,
                                      Ray<T>& r, int k, T fwa, T h, AuxCtx<T>* ctx = nullptr) {
    const T half = T(0.5) * h;
    int cu, cv; T tu, tv;
    double fp[4], ff[4], fa[4];
    // stage 1
    cell_of(r.iu, r.fu, C.nu, cu, tu); cell_of(r.iv, r.fv, C.nv, cv, tv);
    G3<T> g = trilinear<T>(grid, C.nu, C.plane, cu, cv, k, tu, tv, fwa);
    T inv = C.hw / r.dw;
    if (AUX) aux_rates<T>(*ctx, grid, C, cu, cv, k, tu, tv, fwa, r.du, r.dv, r.dw, inv, fp[0], ff[0], fa[0]);
    T aU = C.ru * r.du / r.dw, aV = C.rv * r.dv / r.dw;
    T adu = g.x * inv, adv = g.y * inv, adw = g.z * inv, as = inv;
    bool ok = r.dw > T(0);
    // stage 2
    T du = tfma(half, adu, r.du), dv = tfma(half, adv, r.dv), dw = tfma(half, adw, r.dw);
    cell_of(r.iu, tfma(half, aU, r.fu), C.nu, cu, tu); cell_of(r.iv, tfma(half, aV, r.fv), C.nv, cv, tv);
    g = trilinear<T>(grid, C.nu, C.plane, cu, cv, k, tu, tv, fwa + half);
    ok = ok && dw > T(0);
    inv = C.hw / dw;
    if (AUX) aux_rates<T>(*ctx, grid, C, cu, cv, k, tu, tv, fwa + half, du, dv, dw, inv, fp[1], ff[1], fa[1]);
    T bU = C.ru * du / dw, bV = C.rv * dv / dw;
    T bdu = g.x * inv, bdv = g.y * inv, bdw = g.z * inv, bs = inv;
    // stage 3
    du = tfma(half, bdu, r.du); dv = tfma(half, bdv, r.dv); dw = tfma(half, bdw, r.dw);
    cell_of(r.iu, tfma(half, bU, r.fu), C.nu, cu, tu); cell_of(r.iv, tfma(half, bV, r.fv), C.nv, cv, tv);
    g = trilinear<T>(grid, C.nu, C.plane, cu, cv, k, tu, tv, fwa + half);
    ok = ok && dw > T(0);
    inv = C.hw / dw;
    if (AUX) aux_rates<T>(*ctx, grid, C, cu, cv, k, tu, tv, fwa + half, du, dv, dw, inv, fp[2], ff[2], fa[2]);
    T cU = C.ru * du / dw, cV = C.rv * dv / dw;
    T cdu = g.x * inv, cdv = g.y * inv, cdw = g.z * inv, cs = inv;
    // stage 4
    du = tfma(h, cdu, r.du); dv = tfma(h, cdv, r.dv); dw = tfma(h, cdw, r.dw);
    cell_of(r.iu, tfma(h, cU, r.fu), C.nu, cu, tu); cell_of(r.iv, tfma(h, cV, r.fv), C.nv, cv, tv);
    T fwb = fwa + h;
    g = trilinear<T>(grid, C.nu, C.plane, cu, cv, k, tu, tv, fwb > T(1) ? T(1) : fwb);
    ok = ok && dw > T(0);
    inv = C.hw / dw;
    if (AUX) aux_rates<T>(*ctx, grid, C, cu, cv, k, tu, tv, fwb > T(1) ? T(1) : fwb, du, dv, dw, inv, fp[3], ff[3], fa[3]);
    T eU = C.ru * du / dw, eV = C.rv * dv / dw;
    T edu = g.x * inv, edv = g.y * inv, edw = g.z * inv, es = inv;
    const T h6 = h * T(1.0 / 6.0);
    r.fu = tfma(h6, aU + T(2) * (bU + cU) + eU, r.fu);
    r.fv = tfma(h6, aV + T(2) * (bV + cV) + eV, r.fv);
    r.du = tfma(h6, adu + T(2) * (bdu + cdu) + edu, r.du);
    r.dv = tfma(h6, adv + T(2) * (bdv + cdv) + edv, r.dv);
    r.dw = tfma(h6, adw + T(2) * (bdw + cdw) + edw, r.dw);
    r.s = tfma(h6, as + T(2) * (bs + cs) + es, r.s);
    if (AUX) {
        const double w6 = (double)h6;
        ctx->phase += w6 * (fp[0] + 2.0 * (fp[1] + fp[2]) + fp[3]);
        ctx->farad += w6 * (ff[0] + 2.0 * (ff[1] + ff[2]) + ff[3]);
        ctx->absorb += w6 * (fa[0] + 2.0 * (fa[1] + fa[2]) + fa[3]);
    }
    return ok;
}

// One RK4 step of length ds in path time (general direction).  Fractions un-normalised after.
template <typename T, bool AUX = false>
TT_HD void sstep(const typename GridT<T>::V4* __restrict__ grid, const Consts<T>& C,
                                      Ray<T>& r, T ds, AuxCtx<T>* ctx = nullptr) {
    const T half = T(0.5) * ds;
    int cu, cv, cw; T tu, tv, tw;
    double fp[4], ff[4], fa[4];
    int stage = 0;
    auto field = [&](T fu, T fv, T fw, T du, T dv, T dw) {
        cell_of(r.iu, fu, C.nu, cu, tu); cell_of(r.iv, fv, C.nv, cv, tv); cell_of(r.iw, fw, C.nw, cw, tw);
        if (AUX) { aux_rates<T>(*ctx, grid, C, cu, cv, cw, tu, tv, tw, du, dv, dw, T(1), fp[stage], ff[stage], fa[stage]); ++stage; }
        return trilinear<T>(grid, C.nu, C.plane, cu, cv, cw, tu, tv, tw);
    };
    G3<T> g1 = field(r.fu, r.fv, r.fw, r.du, r.dv, r.dw);
    T d1u = r.du, d1v = r.dv, d1w = r.dw;
    T d2u = tfma(half, g1.x, r.du), d2v = tfma(half, g1.y, r.dv), d2w = tfma(half, g1.z, r.dw);
    G3<T> g2 = field(tfma(half * C.iu_, d1u, r.fu), tfma(half * C.iv_, d1v, r.fv), tfma(half * C.iw_, d1w, r.fw), d2u, d2v, d2w);
    T d3u = tfma(half, g2.x, r.du), d3v = tfma(half, g2.y, r.dv), d3w = tfma(half, g2.z, r.dw);
    G3<T> g3 = field(tfma(half * C.iu_, d2u, r.fu), tfma(half * C.iv_, d2v, r.fv), tfma(half * C.iw_, d2w, r.fw), d3u, d3v, d3w);
    T d4u = tfma(ds, g3.x, r.du), d4v = tfma(ds, g3.y, r.dv), d4w = tfma(ds, g3.z, r.dw);
    G3<T> g4 = field(tfma(ds * C.iu_, d3u, r.fu), tfma(ds * C.iv_, d3v, r.fv), tfma(ds * C.iw_, d3w, r.fw), d4u, d4v, d4w);
    const T s6 = ds * T(1.0 / 6.0);
    if (AUX) {
        const double w6 = (double)s6;
        ctx->phase += w6 * (fp[0] + 2.0 * (fp[1] + fp[2]) + fp[3]);
        ctx->farad += w6 * (ff[0] + 2.0 * (ff[1] + ff[2]) + ff[3]);
        ctx->absorb += w6 * (fa[0] + 2.0 * (fa[1] + fa[2]) + fa[3]);
    }
    r.fu = tfma(s6 * C.iu_, d1u + T(2) * (d2u + d3u) + d4u, r.fu);
    r.fv = tfma(s6 * C.iv_, d1v + T(2) * (d2v + d3v) + d4v, r.fv);
    r.fw = tfma(s6 * C.iw_, d1w + T(2) * (d2w + d3w) + d4w, r.fw);
    r.du = tfma(s6, g1.x + T(2) * (g2.x + g3.x) + g4.x, r.du);
    r.dv = tfma(s6, g1.y + T(2) * (g2.y + g3.y) + g4.y, r.dv);
    r.dw = tfma(s6, g1.z + T(2) * (g2.z + g3.z) + g4.z, r.dw);
    r.s += ds;
}

// fraction of the chord old -> new at which coordinate (i + f) leaves [0, n-1]; 2 if it does not
template <typename T>
TT_HD T leave_fraction(int i, T f_old, T f_new, int n) {
    T lo = T(-i), hi = T(n - 1 - i);
    if (f_new < lo) return (lo - f_old) / (f_new - f_old);
    if (f_new > hi) return (hi - f_old) / (f_new - f_old);
    return T(2);
}

template <typename T>
TT_HD void renorm(int& i, T& f) {
    T fl = tfloor(f);
    i += (int)fl;
    f -= fl;
}
// snap a coordinate that should lie on/inside the faces back into [0, n-1]
template <typename T>
TT_HD void clamp_in(int& i, T& f, int n) {
    renorm(i, f);
    if (i < 0) { i = 0; f = T(0); }
    if (i > n - 1 || (i == n - 1 && f > T(0))) { i = n - 1; f = T(0); }
}


template <typename T> TT_HD T trcp(T x);
#ifndef TT_RCP_NEWTON
#define TT_RCP_NEWTON 0     // measured on B200 (513^3, 1e8 rays): 442.5 ms with the Newton step, 426.8 ms without; the
                            // 2^-23 relative error of MUFU.RCP moves exit positions by ~1e-11 m (tolerance 5e-8 m)
#endif
template <> TT_HD float trcp<float>(float x) {
#ifdef __CUDA_ARCH__
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
#if TT_RCP_NEWTON
    return fmaf(r, fmaf(-x, r, 1.0f), r);        // one Newton step: ~1 ulp
#else
    return r;                                    // MUFU.RCP alone: max relative error 2^-23
#endif
#else
    return 1.0f / x;                             // host run of the kernel source (tests/host/)
#endif
}
template <> TT_HD double trcp<double>(double x) { return 1.0 / x; }


// sqrt(x) for x in (0, 1] as x * rsqrt(x): MUFU.RSQ + one multiplication (2 ulp) instead of the IEEE sequence (MUFU.RSQ,
// two Newton steps, a slow-path branch: ~12 instructions, four times per step)
TT_HD float tsqrt01(float x) {
#ifdef __CUDA_ARCH__
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return x * r;
#else
    return x / sqrtf(x);                                // host run of the kernel source (tests/host/)
#endif
}


struct AuxArgs {
    double omega_over_c;    // phase = omega/c * int (n - 1) ds
    double verdet_nc;       // rotation = V * nc * int (ne/nc) (B . d) ds
};


// trace.cu: frame geometry of a launch; the general kernel over the rays an event kernel deferred
int fill_trace_args(TraceArgs& A, const int n_xyz[3], const double origin_xyz[3], const double spacing_xyz[3], int par);
int launch_trace_second_pass(int dtype, const void* grid4, const double* s0, const uint32_t* perm, double* rf, double* sf,
                             unsigned long long* ray_steps, uint8_t* status, const TraceArgs& A, cudaStream_t s);

// event-marching kernels (trace_event.cu).  aux_out != nullptr selects the variant that also carries the
// passive quantities (FP32 only).  packed = FP32x2 arithmetic (FP32 only).
int launch_trace_event(int dtype, bool packed, int steps_per_cell, const void* grid4, const double* s0,
                       const uint32_t* perm, double* rf, double* sf, unsigned long long* ray_steps, uint8_t* status,
                       const TraceArgs& A, const void* aux4, double* aux_out, const AuxArgs& AX, cudaStream_t stream);

}  // namespace tt
