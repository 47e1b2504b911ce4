// TEST HARNESS (not product): runs the per-ray bodies of the uniform-grid event-marching kernels
// (csrc/trace_event_ray.cuh: event_ray -- tt_trace variant 4, and variant 3 in FP64 -- and event_ray_f32x2, the
// packed FP32x2 production kernel, whose PTX instructions are emulated lane by lane) on the HOST for
// tests/test_host_kernels.py.
#include "trace_event_ray.cuh"
#include "trace_gather_ray.cuh"

static void fill(tt::TraceArgs& A, const int n_xyz[3], const double origin_xyz[3], const double spacing_xyz[3], int par,
                 double extent, double s_max, int spc, long np) {
    using namespace tt;
    Frame f = frame_of(par);
    for (int k = 0; k < 3; ++k) {
        A.fa[k] = f.a[k]; A.n[k] = n_xyz[f.a[k]]; A.o[k] = origin_xyz[f.a[k]]; A.h[k] = spacing_xyz[f.a[k]];
    }
    A.any_deferred = nullptr;
    A.plane_elems = (long long)A.n[0] * A.n[1];
    A.hwf = (float)A.h[2]; A.ruf = (float)(A.h[2] / A.h[0]); A.rvf = (float)(A.h[2] / A.h[1]);
    A.extent = extent; A.s_max = s_max; A.spc = spc; A.np = np;
}

// the production kernel's body (packed FP32x2 arithmetic, emulated lane by lane on the host), as launch_trace_event
// dispatches it; aux4 / aux_out non-null: the tt_trace_aux variant with the passive quantities on board
extern "C" int host_trace_event_packed(const void* grid4, const int n_xyz[3], const double origin_xyz[3],
                                       const double spacing_xyz[3], int par, double extent, double s_max, int spc,
                                       const double* s0, long np, double* rf, double* sf, unsigned char* status,
                                       unsigned long long* ray_steps, long* n_deferred, const void* aux4,
                                       double* aux_out, double omega_over_c, double verdet_nc, int with_aux) {
    using namespace tt;
    TraceArgs A;
    fill(A, n_xyz, origin_xyz, spacing_xyz, par, extent, s_max, spc, np);
    AuxArgs AX;
    AX.omega_over_c = omega_over_c; AX.verdet_nc = verdet_nc;
    const bool spc1 = spc == 1, cubic = A.ruf == 1.0f && A.rvf == 1.0f;
    const float4* g = (const float4*)grid4;
    const float4* b = (const float4*)aux4;
    unsigned long long steps = 0;
    long nd = 0;
    for (long ray = 0; ray < np; ++ray) {
        bool d = false;
#define TT_CALL(S1, AX_, CU) event_ray_f32x2<S1, AX_, CU>(g, s0, ray, rf, sf, status, A, b, aux_out, AX, d)
        if (with_aux) {
            if (spc1) steps += cubic ? TT_CALL(true, true, true) : TT_CALL(true, true, false);
            else steps += cubic ? TT_CALL(false, true, true) : TT_CALL(false, true, false);
        } else {
            if (spc1) steps += cubic ? TT_CALL(true, false, true) : TT_CALL(true, false, false);
            else steps += cubic ? TT_CALL(false, false, true) : TT_CALL(false, false, false);
        }
#undef TT_CALL
        nd += d;
    }
    *ray_steps = steps;
    *n_deferred = nd;
    return 0;
}

extern "C" int host_trace_event(const void* grid4, int dtype, const int n_xyz[3], const double origin_xyz[3],
                                const double spacing_xyz[3], int par, double extent, double s_max, int spc,
                                const double* s0, long np, double* rf, double* sf, unsigned char* status,
                                unsigned long long* ray_steps, long* n_deferred) {
    using namespace tt;
    TraceArgs A;
    fill(A, n_xyz, origin_xyz, spacing_xyz, par, extent, s_max, spc, np);
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

// tt_trace as the library dispatches it, on the host: variant 0 / 3 = event marching (packed FP32x2 body for float
// grids, scalar body for double) followed by the gather kernel's body on the rays it deferred; 1 = 8-corner gather at
// every stage; 2 = cell cache.  with_aux: the tt_trace_aux entry (gather variant 1 with the passive quantities).
extern "C" int host_trace(const void* grid4, int dtype, int variant, const int n_xyz[3], const double origin_xyz[3],
                          const double spacing_xyz[3], int par, double extent, double s_max, int spc,
                          const double* s0, long np, double* rf, double* sf, unsigned char* status,
                          unsigned long long* ray_steps, long* n_deferred) {
    using namespace tt;
    TraceArgs A;
    fill(A, n_xyz, origin_xyz, spacing_xyz, par, extent, s_max, spc, np);
    const AuxArgs AX = AuxArgs();
    const bool spc1 = spc == 1, cubic = A.ruf == 1.0f && A.rvf == 1.0f;
    unsigned long long steps = 0;
    long nd = 0;
    for (long ray = 0; ray < np; ++ray) {
        bool d = false;
        if (variant == 0 || variant == 3) {
            if (dtype == TT_F32) {
                const float4* g = (const float4*)grid4;
#define TT_CALL(S1, CU) event_ray_f32x2<S1, false, CU>(g, s0, ray, rf, sf, status, A, nullptr, nullptr, AX, d)
                if (spc1) steps += cubic ? TT_CALL(true, true) : TT_CALL(true, false);
                else steps += cubic ? TT_CALL(false, true) : TT_CALL(false, false);
#undef TT_CALL
            } else {
                steps += spc1 ? event_ray<double, true>((const double4*)grid4, s0, ray, rf, sf, status, A, d)
                              : event_ray<double, false>((const double4*)grid4, s0, ray, rf, sf, status, A, d);
            }
            if (!d) continue;
            ++nd;
            if (dtype == TT_F32) steps += gather_ray<float, 0, false>((const float4*)grid4, s0, ray, rf, sf, status, A, nullptr, nullptr, AX);
            else steps += gather_ray<double, 0, false>((const double4*)grid4, s0, ray, rf, sf, status, A, nullptr, nullptr, AX);
        } else if (variant == 1) {
            if (dtype == TT_F32) steps += gather_ray<float, 1, false>((const float4*)grid4, s0, ray, rf, sf, status, A, nullptr, nullptr, AX);
            else steps += gather_ray<double, 1, false>((const double4*)grid4, s0, ray, rf, sf, status, A, nullptr, nullptr, AX);
        } else {
            if (dtype == TT_F32) steps += gather_ray<float, 0, false>((const float4*)grid4, s0, ray, rf, sf, status, A, nullptr, nullptr, AX);
            else steps += gather_ray<double, 0, false>((const double4*)grid4, s0, ray, rf, sf, status, A, nullptr, nullptr, AX);
        }
    }
    *ray_steps = steps;
    *n_deferred = nd;
    return 0;
}

// tt_dndr (ElectronCube.dndr): pos, out: (3, npts) in xyz row order
extern "C" int host_dndr(const void* grid4, int dtype, const int n_xyz[3], const double origin_xyz[3],
                         const double spacing_xyz[3], int par, const double* pos, long npts, double* out) {
    using namespace tt;
    TraceArgs A;
    fill(A, n_xyz, origin_xyz, spacing_xyz, par, 0.0, 0.0, 1, npts);
    for (long i = 0; i < npts; ++i) {
        if (dtype == TT_F32) dndr_point<float>((const float4*)grid4, A, pos, npts, i, out);
        else dndr_point<double>((const double4*)grid4, A, pos, npts, i, out);
    }
    return 0;
}
