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
TT_HD int bin_of(double x, const double* __restrict__ e, int nb) {
    const double lo = ldg_f64(e), hi = ldg_f64(e + nb);
    if (!(x >= lo && x <= hi)) return -1;
    int b = (int)((x - lo) * ((double)nb / (hi - lo)));
    b = b < 0 ? 0 : (b > nb - 1 ? nb - 1 : b);
    while (b > 0 && x < ldg_f64(e + b)) --b;
    while (b < nb - 1 && x >= ldg_f64(e + b + 1)) ++b;
    return b;
}

// r^2 exactly as numpy evaluates r[0]**2 + r[2]**2 (two rounded products, one rounded sum)
TT_HD double radius2(double x, double y) {
    return add_rn(mul_rn(x, x), mul_rn(y, y));
}

TT_HD void apply_program(const OpticsArgs& A, double& x, double& th, double& y, double& ph) {
#ifdef __CUDA_ARCH__
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
#else
    const double nan = __builtin_nan("");
#endif
    for (int i = 0; i < A.n_ops; ++i) {
        const tt_optic o = A.ops[i];
        bool reject = false;
        switch (o.op) {
            case TT_OP_DISTANCE:        // [[1, d], [0, 1]] on (x, theta) and (y, phi)
                x = fma(o.a, th, x);
                y = fma(o.a, ph, y);
                break;
            case TT_OP_LENS:            // [[1, 0], [-1/f, 1]]
                th = add_rn(mul_rn(-1.0 / o.a, x), th);
                ph = add_rn(mul_rn(-1.0 / o.b, y), ph);
                break;
            case TT_OP_CIRC_APERTURE:
                reject = radius2(x, y) > mul_rn(o.a, o.a);
                break;
            case TT_OP_CIRC_STOP:
                reject = radius2(x, y) < mul_rn(o.a, o.a);
                break;
            case TT_OP_ANNULAR_STOP: {
                const double rr = radius2(x, y);
                reject = rr > mul_rn(o.a, o.a) && rr < mul_rn(o.b, o.b);
                break;
            }
            case TT_OP_RECT_APERTURE:   // rejects only rays outside in BOTH axes (:132-135)
                reject = mul_rn(x, x) > mul_rn(o.a, o.a) && mul_rn(y, y) > mul_rn(o.b, o.b);
                break;
            case TT_OP_KNIFE_EDGE: {
                const double c = (o.b == 1.0 || o.b == -1.0) ? x : y;
                reject = o.b > 0 ? (c > o.a) : (c < o.a);
                break;
            }
            default: break;
        }
        // a 4x4 matmul spreads a NaN over all four rows; rejected rays are NaN columns (:78)
        if (reject || x != x || th != th || y != y || ph != ph) { x = th = y = ph = nan; }
    }
}

}  // namespace tt
