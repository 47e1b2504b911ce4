// Per-ray part of the optics + histogram kernel (K5 + K6): the element program applied to one ray in registers and
// numpy.histogram2d's bin search.  Host + device: optics_hist.cu builds the kernel from it, tests/host/optics_host.cu
// runs the same source on the CPU against the reference's fixtures.
#pragma once
#include "common.cuh"

namespace tt {

struct OpticsArgs {
    tt_optic ops[TT_MAX_OPTICS];
    int n_ops;
    double pos_scale;
    int nbx, nby;
    long np;
};

// numpy semantics (searchsorted side='right', last edge inclusive): bin b holds e[b] <= x < e[b+1],
// x == e[nb] goes to bin nb-1, anything else (incl. NaN) is dropped (returns -1).
// LDG: the edges live in global memory (read-only path); otherwise plain loads (shared memory, host).  lo, hi, scale =
// e[0], e[nb], nb / (hi - lo): ray-independent, formed once per thread (the division is ~40 FP64 instructions).  The
// two walks settle on searchsorted's bin wherever the first guess lands.
template <bool LDG>
TT_HD int bin_of_scaled(double x, const double* __restrict__ e, int nb, double lo, double hi, double scale) {
    if (!(x >= lo && x <= hi)) return -1;
    int b = (int)((x - lo) * scale);
    b = b < 0 ? 0 : (b > nb - 1 ? nb - 1 : b);
    if (x < (LDG ? ldg_f64(e + b) : e[b])) {                 // (then b >= 1: x >= e[0]) guess too high: walk down
        do --b; while (b > 0 && x < (LDG ? ldg_f64(e + b) : e[b]));
    } else {
        while (b < nb - 1 && x >= (LDG ? ldg_f64(e + b + 1) : e[b + 1])) ++b;
    }
    return b;
}
TT_HD int bin_of(double x, const double* __restrict__ e, int nb) {
    const double lo = ldg_f64(e), hi = ldg_f64(e + nb);
    return bin_of_scaled<true>(x, e, nb, lo, hi, (double)nb / (hi - lo));
}

// r^2 exactly as numpy evaluates r[0]**2 + r[2]**2 (two rounded products, one rounded sum)
TT_HD double radius2(double x, double y) {
    return add_rn(mul_rn(x, x), mul_rn(y, y));
}

// The per-ray constants of the program, formed ONCE on the host before the launch instead of per ray on the device (an
// FP64 division is ~40 instructions; the two lenses of a 4f relay cost four of them per ray: the detector kernel was
// bound by them, 2.1 ms for 1e8 rays whatever the binning strategy).  IEEE division and multiplication round the same
// way on the host and on the device, so the results are bit for bit those of the per-ray form.
//   lens: a, b := -1/f_x, -1/f_y          apertures / stops: a, b := a^2, b^2
inline void prepare_program(OpticsArgs& A) {
    for (int i = 0; i < A.n_ops; ++i) {
        tt_optic& o = A.ops[i];
        switch (o.op) {
            case TT_OP_LENS: o.a = -1.0 / o.a; o.b = -1.0 / o.b; break;
            case TT_OP_CIRC_APERTURE: case TT_OP_CIRC_STOP: case TT_OP_ANNULAR_STOP: case TT_OP_RECT_APERTURE: {
                volatile double a2 = o.a * o.a, b2 = o.b * o.b;       // (volatile: one rounded product each, no contraction)
                o.a = a2; o.b = b2;
                break;
            }
            default: break;
        }
    }
}

// A: a program that went through prepare_program().  N rays advance through the program side by side: one dispatch of
// the element per N rays and N independent dependency chains (the FP64 latencies of one ray's chain, ~100 dependent
// instructions, were what bound the detector kernel once the divisions were gone).
// NaN rule (a 4x4 matmul spreads a NaN over all four rows; rejected rays are NaN columns, :78): a rejected ray is only
// MARKED while the program runs and becomes a NaN column at the end, together with every ray that carries a NaN in any
// row -- a NaN never turns into a number again under the elements' fma / mul / add, and what a marked ray does at later
// elements does not matter, so the result is that of filling the column after every element (8 selects per element
// and ray: more than the elements' own arithmetic).
#define TT_FOR_RAYS _Pragma("unroll") for (int r = 0; r < N; ++r)
// the elements, rejected rays marked in dead[] (no NaN columns yet)
template <int N>
TT_HD void run_program_n(const OpticsArgs& A, double (&x)[N], double (&th)[N], double (&y)[N], double (&ph)[N], bool (&dead)[N]) {
    TT_FOR_RAYS dead[r] = false;
    for (int i = 0; i < A.n_ops; ++i) {
        const tt_optic o = A.ops[i];
        switch (o.op) {
            case TT_OP_DISTANCE:        // [[1, d], [0, 1]] on (x, theta) and (y, phi)
                TT_FOR_RAYS {
                    x[r] = fma(o.a, th[r], x[r]);
                    y[r] = fma(o.a, ph[r], y[r]);
                }
                break;
            case TT_OP_LENS:            // [[1, 0], [-1/f, 1]]; a, b = -1/f_x, -1/f_y
                TT_FOR_RAYS {
                    th[r] = add_rn(mul_rn(o.a, x[r]), th[r]);
                    ph[r] = add_rn(mul_rn(o.b, y[r]), ph[r]);
                }
                break;
            case TT_OP_CIRC_APERTURE:   // a = R^2
                TT_FOR_RAYS dead[r] = dead[r] || radius2(x[r], y[r]) > o.a;
                break;
            case TT_OP_CIRC_STOP:
                TT_FOR_RAYS dead[r] = dead[r] || radius2(x[r], y[r]) < o.a;
                break;
            case TT_OP_ANNULAR_STOP:
                TT_FOR_RAYS {
                    const double rr = radius2(x[r], y[r]);
                    dead[r] = dead[r] || (rr > o.a && rr < o.b);
                }
                break;
            case TT_OP_RECT_APERTURE:   // rejects only rays outside in BOTH axes (:132-135); a, b = Lx^2, Ly^2
                TT_FOR_RAYS dead[r] = dead[r] || (mul_rn(x[r], x[r]) > o.a && mul_rn(y[r], y[r]) > o.b);
                break;
            case TT_OP_KNIFE_EDGE: {
                const bool along_x = o.b == 1.0 || o.b == -1.0, above = o.b > 0;
                TT_FOR_RAYS {
                    const double c = along_x ? x[r] : y[r];
                    dead[r] = dead[r] || (above ? (c > o.a) : (c < o.a));
                }
                break;
            }
            default: break;
        }
    }
}

template <int N>
TT_HD void apply_program_n(const OpticsArgs& A, double (&x)[N], double (&th)[N], double (&y)[N], double (&ph)[N]) {
#ifdef __CUDA_ARCH__
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
#else
    const double nan = __builtin_nan("");
#endif
    bool dead[N];
    run_program_n<N>(A, x, th, y, ph, dead);
    if (A.n_ops > 0) TT_FOR_RAYS {                           // (an empty program leaves the rays as they are)
        if (dead[r] || x[r] != x[r] || th[r] != th[r] || y[r] != y[r] || ph[r] != ph[r]) { x[r] = th[r] = y[r] = ph[r] = nan; }
    }
}
#undef TT_FOR_RAYS

TT_HD void apply_program(const OpticsArgs& A, double& x, double& th, double& y, double& ph) {
    double X[1] = {x}, T[1] = {th}, Y[1] = {y}, P[1] = {ph};
    apply_program_n<1>(A, X, T, Y, P);
    x = X[0]; th = T[0]; y = Y[0]; ph = P[0];
}

}  // namespace tt
