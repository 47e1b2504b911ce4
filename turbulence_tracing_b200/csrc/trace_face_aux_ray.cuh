// Event marching over the face-coefficient grids WITH the passive quantities of tt_trace_aux on board (phase, Faraday
// rotation, inverse-bremsstrahlung attenuation; BASELINE configs[3], SURVEY 8f1 -- the reference holds only the call
// sites: example_kitchensink.py:72-101).
//
// The packed corner-grid kernel (trace_event_ray.cuh, AUX = true) rebuilds four packed trilinear polynomials from 8 corners
// of two node grids per ray and plane (~140 packed operations per step before any evaluation).  Here the bilinear
// coefficients of the five passive fields are formed ONCE per cell face by face_aux_cell() -- a second face grid next to the
// one of trace_face_ray.cuh, five 16-byte words per face cell:
//      word 0 = (A, C, B, D) of ne/nc
//      word 1 = (A_bu, A_bv, B_bu, B_bv)   word 2 = (C_bu, C_bv, D_bu, D_bv)      B_u h_u/h_w and B_v h_v/h_w: with the SCALED
//      word 3 = (A_bw, A_k,  B_bw, B_k )   word 4 = (C_bw, C_k,  D_bw, D_k )      direction e of the face kernel, B.d = b.e
// in the centred cell coordinates of the gradient faces.  The ray loop is the "rebase" form of the face kernel (every step
// runs the whole-cell arithmetic on the polynomial of its own w-interval; a lane that stops at a side face or starts inside
// a cell rebases its polynomials once): one instantiation of the step, no warp vote.  The RK4 stages double as Simpson nodes
// of the three line integrals, summed per step in FP32 and accumulated in FP64, exactly as in the corner-grid kernel; the
// integrands are taken per unit w-fraction WITHOUT the constant h_w, which multiplies the sums at the end.
// Compiled for the device (trace_face_aux.cu) and for the host (tests/host/trace_face_aux_host.cu).
#pragma once
#include "trace_face_ray.cuh"

namespace tt {

// ---- builder: the five words of one face cell from the node grids -----------------------------------------------------
// grid: (g_u, g_v, g_w, ne/nc) nodes; aux: (B_u, B_v, B_w, kappa) nodes or nullptr (phase only: zeros)
TT_HD void face_aux_cell(const float4* __restrict__ grid, const float4* __restrict__ aux, int nu, long long plane, int cu, int cv,
                         int k, double su, double sv, float4* __restrict__ out) {
    const size_t o = (size_t)k * plane + (size_t)cv * nu + cu;
    const float4* p = grid + o;
    const float n00 = GridT<float>::ld(p).w, n10 = GridT<float>::ld(p + 1).w, n01 = GridT<float>::ld(p + nu).w, n11 = GridT<float>::ld(p + nu + 1).w;
    float4 b00 = make_float4(0.f, 0.f, 0.f, 0.f), b10 = b00, b01 = b00, b11 = b00;
    if (aux) {
        const float4* q = aux + o;
        b00 = GridT<float>::ld(q); b10 = GridT<float>::ld(q + 1); b01 = GridT<float>::ld(q + nu); b11 = GridT<float>::ld(q + nu + 1);
    }
#define TT_FA(s, c00, c10, c01, c11, a, b, c, d)                                                                          \
    const float a = (float)(s * (0.25 * (((double)c00 + (double)c10) + ((double)c01 + (double)c11)))),                     \
                b = (float)(s * (0.5 * (((double)c10 - (double)c00) + ((double)c11 - (double)c01)))),                      \
                c = (float)(s * (0.5 * (((double)c01 - (double)c00) + ((double)c11 - (double)c10)))),                      \
                d = (float)(s * ((((double)c11 - (double)c01) - (double)c10) + (double)c00));
    TT_FA(1.0, n00, n10, n01, n11, an, bn, cn, dn)
    TT_FA(su, b00.x, b10.x, b01.x, b11.x, ax, bx, cx, dx)
    TT_FA(sv, b00.y, b10.y, b01.y, b11.y, ay, by, cy, dy)
    TT_FA(1.0, b00.z, b10.z, b01.z, b11.z, az, bz, cz, dz)
    TT_FA(1.0, b00.w, b10.w, b01.w, b11.w, ak, bk, ck, dk)
#undef TT_FA
    out[0] = make_float4(an, cn, bn, dn);
    out[1] = make_float4(ax, ay, bx, by);
    out[2] = make_float4(cx, cy, dx, dy);
    out[3] = make_float4(az, ak, bz, bk);
    out[4] = make_float4(cz, ck, dz, dk);
}
// B_u, B_v are stored times h_u/h_w, h_v/h_w (the inverse of the FP32 ratios the kernel scales the direction with)
inline void face_aux_scales(const double h[3], double& su, double& sv) {
    su = 1.0 / (double)(float)(h[2] / h[0]); sv = 1.0 / (double)(float)(h[2] / h[1]);
}

// ---- the passive polynomials of one face / one w-interval in packed registers ----------------------------------------
struct FaceA {
    f32x2 nac, nbd;                // ne/nc: (A, C), (B, D)
    f32x2 xa, xb, xc, xd;          // (b_u, b_v) lanes
    f32x2 za, zb, zc, zd;          // (b_w, kappa) lanes
};
struct FaceAW {
    float4 w0, w1, w2, w3, w4;
};
TT_HD FaceAW face_a_ldw(const float4* __restrict__ p) {
    FaceAW w;
    w.w0 = GridT<float>::ld(p); w.w1 = GridT<float>::ld(p + 1); w.w2 = GridT<float>::ld(p + 2); w.w3 = GridT<float>::ld(p + 3);
    w.w4 = GridT<float>::ld(p + 4);
    return w;
}
TT_HD FaceA face_a_pack(const FaceAW& w) {
    FaceA q;
    q.nac = pk2(w.w0.x, w.w0.y); q.nbd = pk2(w.w0.z, w.w0.w);
    q.xa = pk2(w.w1.x, w.w1.y); q.xb = pk2(w.w1.z, w.w1.w); q.xc = pk2(w.w2.x, w.w2.y); q.xd = pk2(w.w2.z, w.w2.w);
    q.za = pk2(w.w3.x, w.w3.y); q.zb = pk2(w.w3.z, w.w3.w); q.zc = pk2(w.w4.x, w.w4.y); q.zd = pk2(w.w4.z, w.w4.w);
    return q;
}
#define TT_FA_EACH(OP)                                                                                                   \
    OP(nac) OP(nbd) OP(xa) OP(xb) OP(xc) OP(xd) OP(za) OP(zb) OP(zc) OP(zd)
TT_HD FaceA face_a_sub(const FaceA& f, const FaceA& b) {
    FaceA r;
#define TT_OP(m) r.m = sub2(f.m, b.m);
    TT_FA_EACH(TT_OP)
#undef TT_OP
    return r;
}
TT_HD FaceA face_a_add(const FaceA& f, const FaceA& b) {
    FaceA r;
#define TT_OP(m) r.m = add2(f.m, b.m);
    TT_FA_EACH(TT_OP)
#undef TT_OP
    return r;
}
TT_HD FaceA face_a_scale(const FaceA& f, f32x2 C) {
    FaceA r;
#define TT_OP(m) r.m = mul2(f.m, C);
    TT_FA_EACH(TT_OP)
#undef TT_OP
    return r;
}
TT_HD FaceA face_a_at(const FaceA& base, const FaceA& primed, f32x2 FW) {
    FaceA r;
#define TT_OP(m) r.m = fma2(FW, primed.m, base.m);
    TT_FA_EACH(TT_OP)
#undef TT_OP
    return r;
}
#undef TT_FA_EACH

// the three integrands at one stage, per unit w-fraction and WITHOUT h_w (q = 1 / e_w), ADDED with Simpson weight w to
// the running sums of the step
TT_HD void face_a_integrands(const FaceA& a, f32x2 tuv, f32x2 euv, float ew, float q, float w, float& sp, float& sf_, float& sa) {
    const f32x2 TU = bc2(lo2(tuv)), TV = bc2(hi2(tuv));
    const f32x2 r = fma2(TU, a.nbd, a.nac);
    const float nn = fmaf(hi2(tuv), hi2(r), lo2(r));                                    // ne/nc
    const f32x2 bxy = fma2(TV, fma2(TU, a.xd, a.xc), fma2(TU, a.xb, a.xa));            // (b_u, b_v)
    const f32x2 bzk = fma2(TV, fma2(TU, a.zd, a.zc), fma2(TU, a.zb, a.za));            // (b_w, kappa)
    const float root = tsqrt01(fmaxf(1.f - nn, 1e-30f));
    const float wq = w * q;
    sp = fmaf(-nn * trcp<float>(1.f + root), wq, sp);   // (sqrt(1 - ne/nc) - 1) ds, without cancellation
    const float bd = fmaf(lo2(bxy), lo2(euv), fmaf(hi2(bxy), hi2(euv), lo2(bzk) * ew));
    sf_ = fmaf(nn * bd, wq, sf_);                       // (ne/nc) (B . d) ds
    sa = fmaf(hi2(bzk), wq, sa);                        // kappa ds
}

// Returns the plane arrivals of this ray (0 if it is deferred to the general kernel).  faces / facesA: the coefficient
// grids of tt_build_face_grid / tt_build_face_aux_grid for the same node grids and probing direction.
template <bool TRACK_S>
TT_HD unsigned face_aux_ray_f32x2(const float4* __restrict__ faces, const float4* __restrict__ facesA, const double* __restrict__ s0,
                                  long ray, double* __restrict__ rf, double* __restrict__ sf, double* __restrict__ aux_out,
                                  uint8_t* __restrict__ status, const TraceArgs& A, const FaceArgs& FA, const AuxArgs& AX,
                                  bool& deferred) {
    unsigned steps = 0;
    const int nu = A.n[0], nv = A.n[1], nw = A.n[2];
    const long long cplane = (long long)(nu - 1) * (nv - 1);          // face cells per plane
    const int crow = nu - 1;
    // ---- prologue (the same entry conditions as face_ray_f32x2) ------------------------------------------------
    double X[3], D[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        X[k] = (s0[(size_t)A.fa[k] * A.np + ray] - A.o[k]) / A.h[k];
        D[k] = s0[(size_t)(3 + A.fa[k]) * A.np + ray] * (1.0 / kC);
    }
    double s_pre = 0.0;
    if (X[2] < 0.0 && D[2] > TT_MARCH_MIN_DW) {           // launched in front of the cube: free flight to the entry face
        s_pre = -X[2] * A.h[2] / D[2];
        X[0] += D[0] / A.h[0] * s_pre;
        X[1] += D[1] / A.h[1] * s_pre;
        X[2] = 0.0;
    }
    bool fast = X[0] >= 0.0 && X[0] <= (double)(nu - 1) && X[1] >= 0.0 && X[1] <= (double)(nv - 1) &&
                X[2] >= 0.0 && X[2] <= (double)(nw - 1) && D[2] > TT_MARCH_MIN_DW;
    fast = fast && ((double)(nw - 1) - X[2]) * A.h[2] <= TT_MARCH_MIN_DW * (A.s_max - s_pre);
    int cu = 0, cv = 0, k = 0;
    float tu0 = 0.f, tv0 = 0.f, fw = 0.f;
    if (fast) {
        double fl;
        fl = fmin(floor(X[0]), (double)(nu - 2)); cu = (int)fl; tu0 = (float)((X[0] - fl) - 0.5);     // centred
        fl = fmin(floor(X[1]), (double)(nv - 2)); cv = (int)fl; tv0 = (float)((X[1] - fl) - 0.5);
        fl = floor(X[2]); k = (int)fl; fw = (float)(X[2] - fl);
    }
    f32x2 tuv = pk2(tu0, tv0), duv = pk2((float)D[0] * FA.ruf, (float)D[1] * FA.rvf);     // scaled transverse direction
    float dw = (float)D[2], s = 0.f;
    const bool track_s = TRACK_S && sf != nullptr;
    double acc_p = 0.0, acc_f = 0.0, acc_a = 0.0;         // line integrals of (n - 1), (ne/nc)(B.d), kappa per unit h_w

    if (fast && k < nw - 1) {
        long long cell = (long long)k * cplane + (long long)cv * crow + cu;
        const float4* p = faces + 3 * cell;
        const float4* pa = facesA + 5 * cell;
        FaceQ B = face_ld(p), P = face_sub(face_ld(p + 3 * cplane), B);
        FaceA BA = face_a_pack(face_a_ldw(pa)), PA = face_a_sub(face_a_pack(face_a_ldw(pa + 5 * cplane)), BA);
        // (B, P), (BA, PA) = the polynomials at the start of the step and their change over the step's w-interval
        if (fw != 0.f) {                                  // launched inside a cell
            const f32x2 FW = bc2(fw), R = bc2(1.f - fw);
            B = face_at(B, P, FW); P = face_scale(P, R);
            BA = face_a_at(BA, PA, FW); PA = face_a_scale(PA, R);
        }
        FaceW N;
        FaceAW NA;
        while (true) {
            // faces k+2, consumed when the ray arrives at plane k+1 (unconditional: both grids carry a spare plane)
            N = face_ldw(p + 6 * cplane);
            NA = face_a_ldw(pa + 10 * cplane);
            const float q1 = trcp<float>(dw);             // (dw > TT_MARCH_MIN_DW: checked at the entry and after every step)
            const f32x2 aUV = mul2(duv, bc2(q1));
            float h = 1.f - fw;
            float su = 0.f, sv = 0.f;                     // a predicted side crossing: +-1 on the axis crossed
            {
                const f32x2 puv = fma2(bc2(h), aUV, tuv);
                const float pu = lo2(puv), pv = hi2(puv);
                const bool ou = fabsf(pu) > 0.5f, ov = fabsf(pv) > 0.5f;
                if (ou || ov) {
                    // fraction of the interval at which the chord reaches the face the prediction lies beyond (named by the
                    // sign of the prediction); a ray a rounding error outside gets <= 0: a zero-length step and the relabelling
                    const f32x2 den = mul2(bc2(h), aUV);
                    const float lu = ou ? chord_fraction<float>(copysignf(0.5f, pu) - lo2(tuv), lo2(den)) : 2.f;
                    const float lv = ov ? chord_fraction<float>(copysignf(0.5f, pv) - hi2(tuv), hi2(den)) : 2.f;
                    float lam = fminf(lu, lv);
                    if (lam < 1.f) {
                        if (lu <= lv) su = copysignf(1.f, pu); else sv = copysignf(1.f, pv);
                        lam = fmaxf(lam, 0.f);
                        h *= lam;
                        const f32x2 L = bc2(lam);
                        P = face_scale(P, L); PA = face_a_scale(PA, L);          // the step stops at the side face
                    }
                }
            }
            // ---- stage 1 at the start of the interval ----------------------------------------------------------
            f32x2 g; float gw;
            face_eval(B, tuv, g, gw);
            const f32x2 aduv = mul2(g, bc2(q1));
            const float adw = gw * q1;
            float sp = 0.f, sfar = 0.f, sab = 0.f;        // Simpson sums of the step: 1, 2, 2, 1
            face_a_integrands(BA, tuv, duv, dw, q1, 1.f, sp, sfar, sab);
            const float half = 0.5f * h;
            const f32x2 HALF = bc2(half), H = bc2(h);
            // ---- stages 2 and 3 share the middle of the interval -------------------------------------------------
            f32x2 suv = fma2(HALF, aUV, tuv), duv2 = fma2(HALF, aduv, duv);
            float dw2 = fmaf(half, adw, dw);
            const float q2 = trcp<float>(dw2);
            bool ok = dw2 > 0.f;
            const FaceQ M = face_at(B, P, bc2(0.5f));
            const FaceA MA = face_a_at(BA, PA, bc2(0.5f));
            const f32x2 bUV = mul2(duv2, bc2(q2));
            face_eval(M, suv, g, gw);
            face_a_integrands(MA, suv, duv2, dw2, q2, 2.f, sp, sfar, sab);
            const f32x2 bduv = mul2(g, bc2(q2));
            const float bdw = gw * q2;
            suv = fma2(HALF, bUV, tuv); duv2 = fma2(HALF, bduv, duv); dw2 = fmaf(half, bdw, dw);
            const float q3 = trcp<float>(dw2);
            ok = ok && dw2 > 0.f;
            const f32x2 Q3 = bc2(q3), cUV = mul2(duv2, Q3), sUV = fma2(duv2, Q3, bUV);              // sUV = b + c
            face_eval(M, suv, g, gw);
            face_a_integrands(MA, suv, duv2, dw2, q3, 2.f, sp, sfar, sab);
            const f32x2 cduv = mul2(g, Q3), sduv = fma2(g, Q3, bduv);
            const float cdw = gw * q3, sdw = fmaf(gw, q3, bdw);
            // ---- stage 4 at the end of the interval: B, BA become the polynomials there ------------------------------
            suv = fma2(H, cUV, tuv); duv2 = fma2(H, cduv, duv); dw2 = fmaf(h, cdw, dw);
            const float q4 = trcp<float>(dw2);
            ok = ok && dw2 > 0.f;
            B = face_add(B, P);
            BA = face_a_add(BA, PA);
            face_eval(B, suv, g, gw);
            face_a_integrands(BA, suv, duv2, dw2, q4, 1.f, sp, sfar, sab);
            const float h6 = h * (float)(1.0 / 6.0);
            const f32x2 H6 = bc2(h6), TWO = bc2(2.f), Q4 = bc2(q4);
            tuv = fma2(H6, fma2(duv2, Q4, fma2(TWO, sUV, aUV)), tuv);                                // a + 2 (b + c) + e
            duv = fma2(H6, fma2(g, Q4, fma2(TWO, sduv, aduv)), duv);
            dw = fmaf(h6, fmaf(gw, q4, fmaf(2.f, sdw, adw)), dw);
            if (TRACK_S) s = fmaf(h6, q4 + fmaf(2.f, q2 + q3, q1), s);
            acc_p += (double)(h6 * sp);                                                              // Simpson over the stages
            acc_f += (double)(h6 * sfar);
            acc_a += (double)(h6 * sab);
            if (!(ok && dw > (float)TT_MARCH_MIN_DW)) { fast = false; break; }
            if (su == 0.f && sv == 0.f) {
                ++steps;
                fw = 0.f;
                if (++k >= nw - 1) break;
                p += 3 * cplane; pa += 5 * cplane;
                P = face_sub(face_pack(N), B);                                // B is the far face by now: the next base
                PA = face_a_sub(face_a_pack(NA), BA);
            } else {
                fw += h;
                tuv = sub2(tuv, pk2(su, sv));
                const int du = (int)su, dv = (int)sv;
                cu += du; cv += dv;
                if ((unsigned)cu > (unsigned)(nu - 2) || (unsigned)cv > (unsigned)(nv - 2)) { fast = false; break; }      // side exit
                const int dc = du + dv * crow;
                p += 3 * dc; pa += 5 * dc;
                const f32x2 FW = bc2(fw), R = bc2(1.f - fw);
                const FaceQ Bn = face_ld(p);                                  // both faces of the new cell column
                const FaceQ Pn = face_sub(face_ld(p + 3 * cplane), Bn);
                B = face_at(Bn, Pn, FW);                                      // rebased to the rest of the cell
                P = face_scale(Pn, R);
                const FaceA BAn = face_a_pack(face_a_ldw(pa));
                const FaceA PAn = face_a_sub(face_a_pack(face_a_ldw(pa + 5 * cplane)), BAn);
                BA = face_a_at(BAn, PAn, FW);
                PA = face_a_scale(PAn, R);
            }
        }
    }
    if (!fast) {
        status[ray] = TT_RAY_DEFERRED;
        deferred = true;
        steps = 0;
    } else {
        const double Pu = A.o[0] + (((double)cu + 0.5) + (double)lo2(tuv)) * A.h[0];
        const double Pv = A.o[1] + (((double)cv + 0.5) + (double)hi2(tuv)) * A.h[1];
        const double Pw = A.o[2] + (double)(nw - 1) * A.h[2];
        const double Vu = (double)lo2(duv) * FA.inv_ru * kC, Vv = (double)hi2(duv) * FA.inv_rv * kC, Vw = (double)dw * kC;
        const double tb = (Pw - A.extent) / Vw;
        rf[0 * A.np + ray] = Pu - Vu * tb;
        rf[1 * A.np + ray] = atan(Vu / Vw);
        rf[2 * A.np + ray] = Pv - Vv * tb;
        rf[3 * A.np + ray] = atan(Vv / Vw);
        if (track_s) {
            const double t_rest = (A.s_max - s_pre - (double)s * A.h[2]) / kC;
            const double Pf[3] = {Pu, Pv, Pw}, Vf[3] = {Vu, Vv, Vw};
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                sf[(size_t)A.fa[m] * A.np + ray] = Pf[m] + Vf[m] * t_rest;
                sf[(size_t)(3 + A.fa[m]) * A.np + ray] = Vf[m];
            }
        }
        const double hw = A.h[2];                         // the integrands were taken per unit w-fraction without h_w
        aux_out[0 * A.np + ray] = exp(-0.5 * acc_a * hw);
        aux_out[1 * A.np + ray] = AX.omega_over_c * acc_p * hw;
        aux_out[2 * A.np + ray] = AX.verdet_nc * acc_f * hw;
        status[ray] = (uint8_t)TT_RAY_EXIT_FACE;
    }
    return steps;
}

}  // namespace tt
