// TEST HARNESS (not product): runs the per-ray body of the rectilinear event-marching kernel -- the very source
// the GPU kernel is built from (csrc/trace_axes_event.cuh) -- on the HOST, one ray after the other, so that
// tests/test_host_kernels.py can check it against the C oracle without a GPU.  Built by the test with
//   nvcc -std=c++17 -O1 -shared -Xcompiler -fPIC -I include -I turbulence_tracing_b200/csrc
#include "trace_axes_event.cuh"

extern "C" int host_axes_event(const void* grid4, int dtype, const int n_xyz[3], const double* x, const double* y,
                               const double* z, int par, double extent, double s_max, int spc, const double* s0,
                               long np, double* rf, double* sf, unsigned char* status, unsigned long long* ray_steps,
                               long* n_deferred) {
    using namespace tt;
    AxesArgs A;
    const double* axes[3] = {x, y, z};
    Frame f = frame_of(par);
    for (int k = 0; k < 3; ++k) { A.fa[k] = f.a[k]; A.n[k] = n_xyz[f.a[k]]; A.ax[k] = axes[f.a[k]]; }
    A.extent = extent; A.s_max = s_max; A.spc = spc; A.np = np;
    unsigned long long steps = 0;
    long nd = 0;
    for (long ray = 0; ray < np; ++ray) {
        bool deferred = false;
        if (dtype == TT_F32) steps += axes_event_ray<float>((const float4*)grid4, s0, ray, rf, sf, status, A, deferred);
        else steps += axes_event_ray<double>((const double4*)grid4, s0, ray, rf, sf, status, A, deferred);
        nd += deferred;
    }
    *ray_steps = steps;
    *n_deferred = nd;
    return 0;
}
