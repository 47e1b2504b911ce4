// TEST HARNESS (not product): the per-ray part of the optics + histogram kernel (csrc/optics_program.cuh: element
// program in registers, numpy.histogram2d bin search) run on the HOST, ray after ray, for tests/test_host_kernels.py.
#include "optics_program.cuh"

extern "C" int host_optics_hist(const double* rf_in, long np, double pos_scale, const tt_optic* program, int n_ops,
                                const double* xe, int nbx, const double* ye, int nby, unsigned long long* H,
                                double* rf_out) {
    using namespace tt;
    if (n_ops < 0 || n_ops > TT_MAX_OPTICS) return 1;
    OpticsArgs A;
    for (int i = 0; i < n_ops; ++i) A.ops[i] = program[i];
    A.n_ops = n_ops; A.pos_scale = pos_scale; A.nbx = nbx; A.nby = nby; A.np = np;
    prepare_program(A);
    for (long ray = 0; ray < np; ++ray) {
        double x = rf_in[ray] * A.pos_scale, th = rf_in[np + ray];
        double y = rf_in[2 * np + ray] * A.pos_scale, ph = rf_in[3 * np + ray];
        apply_program(A, x, th, y, ph);
        if (rf_out) { rf_out[ray] = x; rf_out[np + ray] = th; rf_out[2 * np + ray] = y; rf_out[3 * np + ray] = ph; }
        if (H) {
            const int ix = bin_of(x, xe, nbx), iy = bin_of(y, ye, nby);
            if (ix >= 0 && iy >= 0) H[(size_t)iy * nbx + ix] += 1ull;
        }
    }
    return 0;
}
