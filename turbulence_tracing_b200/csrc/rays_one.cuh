// One launch ray of tt_init_beam and one Morton key of tt_sort_rays (see rays.cu): host + device, so that
// tests/host/rays_host.cu can generate beams and keys on the CPU.
#pragma once
#ifndef _GNU_SOURCE
#define _GNU_SOURCE          // sincos on the host
#endif
#include <math.h>
#include "common.cuh"

namespace tt {

static constexpr double kC = 299792458.0;
static constexpr double kPi = 3.14159265358979323846;

// ray `first + i` of the global beam (counter RNG: any shard can be generated anywhere), written to column i of s0[6][np]
TT_HD void init_beam_ray(long i, long np, long first, uint64_t seed, double beam_size, double divergence, double extent,
                         int par, double* __restrict__ s0) {
    const uint64_t id = (uint64_t)(first + i);
    // draws of particle_tracker.py:273-278: t, u1, u2, phi, chi (chi ~ N(0,1) via Box-Muller)
    Philox a = philox4x32_10(id, 0, seed), b = philox4x32_10(id, 1, seed), c = philox4x32_10(id, 2, seed);
    const double t = 2.0 * kPi * u01(a.c[0], a.c[1]);
    double u = u01(a.c[2], a.c[3]) + u01(b.c[0], b.c[1]);
    if (u > 1.0) u = 2.0 - u;
    const double phi = kPi * u01(b.c[2], b.c[3]);
    const double chi = divergence * sqrt(-2.0 * log(u01(c.c[0], c.c[1]))) * cos(2.0 * kPi * u01(c.c[2], c.c[3]));
    double st, ct, sp, cp, sc, cc;
    sincos(t, &st, &ct);
    sincos(phi, &sp, &cp);
    sincos(chi, &sc, &cc);
    const double p1 = beam_size * u * ct, p2 = beam_size * u * st;
    const double vpar = kC * cc, v1 = kC * sc * cp, v2 = kC * sc * sp;
    // transverse axes (t1, t2) and the launch plane; 'x' launches at +extent (quirk of :280-289)
    const Frame f = frame_of(par);
    const double ppar = par == 0 ? extent : -extent;
    s0[(size_t)f.a[0] * np + i] = p1;
    s0[(size_t)f.a[1] * np + i] = p2;
    s0[(size_t)f.a[2] * np + i] = ppar;
    s0[(size_t)(3 + f.a[0]) * np + i] = v1;
    s0[(size_t)(3 + f.a[1]) * np + i] = v2;
    s0[(size_t)(3 + f.a[2]) * np + i] = vpar;
}

TT_HD uint32_t spread16(uint32_t v) {
    v &= 0xFFFFu;
    v = (v | (v << 8)) & 0x00FF00FFu;
    v = (v | (v << 4)) & 0x0F0F0F0Fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}

// Z-order key of ray i: 16 bits per transverse axis over the cube's width
TT_HD uint32_t morton_key(const double* __restrict__ s0, long i, long np, int au, int av, double ou, double ov,
                          double su, double sv) {
    // 16 bits per transverse axis over the cube's width (su, sv = 65536 / width)
    double qu = (s0[(size_t)au * np + i] - ou) * su, qv = (s0[(size_t)av * np + i] - ov) * sv;
    qu = fmin(fmax(qu, 0.0), 65535.0);
    qv = fmin(fmax(qv, 0.0), 65535.0);
    if (!(qu == qu)) qu = 0.0;
    if (!(qv == qv)) qv = 0.0;
    return spread16((uint32_t)qu) | (spread16((uint32_t)qv) << 1);
}

}  // namespace tt
