// TEST HARNESS (not product): runs the per-ray body of the uniform-grid event-marching kernels
// (csrc/trace_event_ray.cuh -- tt_trace variant 4, and variant 3 in FP64; the packed FP32x2 production kernel
// performs the same operations in the same order) on the HOST for tests/test_host_kernels.py.
#include "trace_event_ray.cuh"

extern "C" int host_trace_event(const void* grid4, int dtype, const int n_xyz[3], const double origin_xyz[3],
                                const double spacing_xyz[3], int par, double extent, double s_max, int spc,
                                const double* s0, long np, double* rf, double* sf, unsigned char* status,
                                unsigned long long* ray_steps, long* n_deferred) {
    using namespace tt;
    TraceArgs A;
    Frame f = frame_of(par);
    for (int k = 0; k < 3; ++k) {
        A.fa[k] = f.a[k]; A.n[k] = n_xyz[f.a[k]]; A.o[k] = origin_xyz[f.a[k]]; A.h[k] = spacing_xyz[f.a[k]];
    }
    A.any_deferred = nullptr;
    A.plane_elems = (long long)A.n[0] * A.n[1];
    A.hwf = (float)A.h[2]; A.ruf = (float)(A.h[2] / A.h[0]); A.rvf = (float)(A.h[2] / A.h[1]);
    A.extent = extent; A.s_max = s_max; A.spc = spc; A.np = np;
    unsigned long long steps = 0;
    long nd = 0;
    for (long ray = 0; ray < np; ++ray) {
        bool deferred = false;
        if (dtype == TT_F32) {
            steps += spc == 1 ? event_ray<float, true>((const float4*)grid4, s0, ray, rf, sf, status, A, deferred)
                              : event_ray<float, false>((const float4*)grid4, s0, ray, rf, sf, status, A, deferred);
        } else {
            steps += spc == 1 ? event_ray<double, true>((const double4*)grid4, s0, ray, rf, sf, status, A, deferred)
                              : event_ray<double, false>((const double4*)grid4, s0, ray, rf, sf, status, A, deferred);
        }
        nd += deferred;
    }
    *ray_steps = steps;
    *n_deferred = nd;
    return 0;
}
