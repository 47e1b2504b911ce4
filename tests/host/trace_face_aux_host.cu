// TEST HARNESS (not product): the builders of the two face-coefficient grids and the per-ray body of the face kernel with the
// passive quantities (face_aux_ray_f32x2, csrc/trace_face_aux_ray.cuh; packed FP32x2 instructions emulated lane by lane) on
// the HOST for tests/test_host_kernels.py.
#include "trace_face_aux_ray.cuh"

extern "C" int host_trace_faces_aux(const void* grid4, const void* aux4, const int n_xyz[3], const double origin_xyz[3],
                                    const double spacing_xyz[3], int par, double extent, double s_max, const double* s0, long np,
                                    double* rf, double* sf, double* aux_out, unsigned char* status, unsigned long long* ray_steps,
                                    long* n_deferred, double omega_over_c, double verdet_nc, void* faces, void* facesA) {
    using namespace tt;
    TraceArgs A;
    Frame f = frame_of(par);
    for (int k = 0; k < 3; ++k) {
        A.fa[k] = f.a[k]; A.n[k] = n_xyz[f.a[k]]; A.o[k] = origin_xyz[f.a[k]]; A.h[k] = spacing_xyz[f.a[k]];
    }
    A.any_deferred = nullptr;
    A.plane_elems = (long long)A.n[0] * A.n[1];
    A.hwf = (float)A.h[2]; A.ruf = (float)(A.h[2] / A.h[0]); A.rvf = (float)(A.h[2] / A.h[1]);
    A.extent = extent; A.s_max = s_max; A.spc = 1; A.np = np;
    FaceArgs FA;
    fill_face_args(FA, A);
    AuxArgs AX;
    AX.omega_over_c = omega_over_c; AX.verdet_nc = verdet_nc;
    double su, sv, sw, au, av;
    face_scales(A.h, su, sv, sw);
    face_aux_scales(A.h, au, av);
    const long long plane = (long long)A.n[0] * A.n[1];
    float4* out = (float4*)faces;
    float4* outA = (float4*)facesA;
    for (int kk = 0; kk <= A.n[2]; ++kk)
        for (int cv = 0; cv < A.n[1] - 1; ++cv)
            for (int cu = 0; cu < A.n[0] - 1; ++cu) {
                const int k = kk < A.n[2] ? kk : A.n[2] - 1;
                face_grid_cell((const float4*)grid4, A.n[0], plane, cu, cv, k, su, sv, sw, out);
                face_aux_cell((const float4*)grid4, (const float4*)aux4, A.n[0], plane, cu, cv, k, au, av, outA);
                out += 3; outA += 5;
            }
    unsigned long long steps = 0;
    long nd = 0;
    for (long ray = 0; ray < np; ++ray) {
        bool d = false;
        steps += sf ? face_aux_ray_f32x2<true>((const float4*)faces, (const float4*)facesA, s0, ray, rf, sf, aux_out, status, A, FA, AX, d)
                    : face_aux_ray_f32x2<false>((const float4*)faces, (const float4*)facesA, s0, ray, rf, sf, aux_out, status, A, FA, AX, d);
        if (d) ++nd;
    }
    *ray_steps = steps;
    *n_deferred = nd;
    return 0;
}
