// TEST HARNESS (not product): runs the builder of the face-coefficient grid (face_grid_cell) and the per-ray body of
// the production kernel over it (face_ray_f32x2, csrc/trace_face_ray.cuh; packed FP32x2 instructions emulated lane by
// lane) on the HOST for tests/test_host_kernels.py, followed -- as tt_trace_faces does -- by the gather kernel's body
// on the rays it deferred.
#include "trace_face_ray.cuh"
#include "trace_gather_ray.cuh"

static void fill(tt::TraceArgs& A, const int n_xyz[3], const double origin_xyz[3], const double spacing_xyz[3], int par,
                 double extent, double s_max, long np) {
    using namespace tt;
    Frame f = frame_of(par);
    for (int k = 0; k < 3; ++k) {
        A.fa[k] = f.a[k]; A.n[k] = n_xyz[f.a[k]]; A.o[k] = origin_xyz[f.a[k]]; A.h[k] = spacing_xyz[f.a[k]];
    }
    A.any_deferred = nullptr;
    A.plane_elems = (long long)A.n[0] * A.n[1];
    A.hwf = (float)A.h[2]; A.ruf = (float)(A.h[2] / A.h[0]); A.rvf = (float)(A.h[2] / A.h[1]);
    A.extent = extent; A.s_max = s_max; A.spc = 1; A.np = np;
}

// faces: 48 (nu-1)(nv-1)(nw+1) bytes, as tt_build_face_grid writes them (the last plane repeats face nw-1)
extern "C" int host_build_face_grid(const void* grid4, const int n_xyz[3], const double spacing_xyz[3], int par, void* faces) {
    using namespace tt;
    const Frame f = frame_of(par);
    int n[3];
    double h[3];
    for (int k = 0; k < 3; ++k) { n[k] = n_xyz[f.a[k]]; h[k] = spacing_xyz[f.a[k]]; }
    double su, sv, sw;
    face_scales(h, su, sv, sw);
    const long long plane = (long long)n[0] * n[1];
    float4* out = (float4*)faces;
    for (int kk = 0; kk <= n[2]; ++kk)
        for (int cv = 0; cv < n[1] - 1; ++cv)
            for (int cu = 0; cu < n[0] - 1; ++cu) {
                const int k = kk < n[2] ? kk : n[2] - 1;
                face_grid_cell((const float4*)grid4, n[0], plane, cu, cv, k, su, sv, sw, out);
                out += 3;
            }
    return 0;
}

extern "C" int host_trace_faces(const void* grid4, const void* faces, const int n_xyz[3], const double origin_xyz[3],
                                const double spacing_xyz[3], int par, double extent, double s_max, const double* s0, long np,
                                double* rf, double* sf, unsigned char* status, unsigned long long* ray_steps,
                                long* n_deferred, int second_pass) {
    using namespace tt;
    TraceArgs A;
    fill(A, n_xyz, origin_xyz, spacing_xyz, par, extent, s_max, np);
    FaceArgs FA;
    fill_face_args(FA, A);
    const AuxArgs AX = AuxArgs();
    unsigned long long steps = 0;
    long nd = 0;
    for (long ray = 0; ray < np; ++ray) {
        bool d = false;
        steps += sf ? face_ray_f32x2<true>((const float4*)faces, s0, ray, rf, sf, status, A, FA, d)
                    : face_ray_f32x2<false>((const float4*)faces, s0, ray, rf, sf, status, A, FA, d);
        if (!d) continue;
        ++nd;
        if (second_pass) steps += gather_ray<float, 0, false>((const float4*)grid4, s0, ray, rf, sf, status, A, nullptr, nullptr, AX);
    }
    *ray_steps = steps;
    *n_deferred = nd;
    return 0;
}
