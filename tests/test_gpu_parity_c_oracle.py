"""GPU parity at BASELINE sizes against the C oracle (oracle/tt_oracle.c, pinned to the live reference's
fixtures by tests/test_oracle_c.py).  The scipy oracle needs minutes for a few hundred tight-tolerance rays;
the C restatement gives every ray its own step control (batch=1) at rtol = 1e-13 -- self-converged to better
than 1e-7 of the rms angle -- in milliseconds, so the same criteria can be checked for thousands of rays on the full
257^3 cube of BASELINE configs[1].

Tolerances (BASELINE.json north_star):
  FP64 mode : exit positions within 1e-5 of the beam radius, angles within 1e-5 of the rms angle
  FP32 mode : exit position within 1e-3 of a detector pixel (52 nm at the default bin_scale = 10)
  histogram : L1(H - H_ref) / sum(H_ref) <= 1e-3 (only rays within ~1 nm of a bin edge can differ)
"""
import numpy as np
import pytest

from oracle import c_oracle as orc_c
from oracle import ref_numpy as orc

pytestmark = pytest.mark.gpu

PIXEL_M = 18e-3 / (3448 // 10)
BEAM = 4e-3


@pytest.fixture(scope="module")
def tt():
    import torch
    assert torch.cuda.is_available()
    import turbulence_tracing_b200 as pkg
    from turbulence_tracing_b200 import _lib
    _lib.load(build_if_missing=False)
    return pkg


def _errors(rf, ref):
    pos = np.abs(rf[0::2] - ref[0::2]).max()
    rms = max(np.sqrt(np.mean(ref[1] ** 2 + ref[3] ** 2)), 1e-6)
    return pos, np.abs(rf[1::2] - ref[1::2]).max() / rms, rms


def _trace(tt, x, y, z, ne, s0, extent, direction, dtype, spc):
    cube = tt.particle_tracker.ElectronCube(x, y, z, probing_direction=direction, dtype=dtype, steps_per_cell=spc,
                                            verbose=False)
    cube.external_ne(ne)
    cube.calc_dndr()
    cube.s0 = s0
    cube.extent = float(extent)
    return cube, np.asarray(cube.solve())


def test_c2_cube_sample_against_c_oracle(tt):
    """BASELINE configs[1]: the 257^3 k^-11/3 cube of the device generator, 8192 rays of the configs' beam."""
    tg, rtm = tt.turboGen, tt.ray_transfer_matrix
    f = tg.gaussian3D_FFT(128, lambda k: k ** (-11.0 / 3.0), seed=1234, dtype="float32", return_device=True).torch
    ne = (1e25 * (1 + 0.3 * f / f.std()).clamp(min=0)).cpu().numpy()          # float32 values, identical on both sides
    x = np.linspace(-5e-3, 5e-3, 257)
    np.random.seed(11)
    s0 = orc.init_beam(8192, BEAM, 0.05e-3, 5e-3, "z")
    field = orc_c.make_field(ne.astype(np.float64), x, x, x)
    ref = orc_c.solve(field, s0, 5e-3, "z", rtol=1e-13, atol=1e-16, batch=1)[0]
    ref2 = orc_c.solve(field, s0[:, :512], 5e-3, "z", rtol=1e-12, atol=1e-15, batch=1)[0]
    p, a, rms = _errors(ref2, ref[:, :512])
    print(f"C oracle self-convergence (rtol 1e-12 vs 1e-13): {p:.1e} m, {a:.1e} of the rms angle ({rms * 1e3:.2f} mrad)")
    assert p <= 1e-7 * BEAM and a <= 1e-6

    _, rf64 = _trace(tt, x, x, x, ne, s0, 5e-3, "z", "float64", 4)
    p, a, _ = _errors(rf64, ref)
    print(f"257^3 fp64 4 steps/cell vs C oracle: {p:.2e} m = {p / BEAM:.1e} of the beam radius, angle {a:.1e} of rms")
    assert p <= 1e-5 * BEAM and a <= 1e-5

    cube32, rf32 = _trace(tt, x, x, x, ne, s0, 5e-3, "z", "float32", 1)
    p, a, _ = _errors(rf32, ref)
    print(f"257^3 fp32 1 step/cell (production setting) vs C oracle: {p:.2e} m = {p / PIXEL_M:.1e} pixel, angle {a:.1e} of rms")
    assert p <= 1e-3 * PIXEL_M
    assert cube32.ray_steps == 256 * s0.shape[1]

    # detector images from the two sets of rays
    for name, cls, kw in (("shadowgraphy", rtm.Shadowgraphy, {}), ("schlieren_df", rtm.Schlieren_DF, {"R": 1}),
                          ("schlieren_lf", rtm.Schlieren_LF, {"R": 1})):
        det = cls(cube32.rf)
        det.solve(**kw)
        det.histogram()
        Href, _, _ = orc.histogram(orc.detector(name, ref, **({"R_stop": kw["R"]} if kw else {})))
        l1 = np.abs(det.H - Href).sum() / max(Href.sum(), 1)
        print(f"{name}: {int(det.H.sum())} rays binned, L1 distance to the oracle image {l1:.1e}")
        assert l1 <= 1e-3


def test_non_cubic_cells_probing_y_against_c_oracle(tt):
    """Cells of three different sizes (97 x 129 x 65 nodes over the same extent -> 104 / 78 / 156 um), beam along
    y: the index-space slope rescaling of the event kernels (h_w/h_u, h_w/h_v != 1) against the C oracle."""
    x, y, z = np.linspace(-5e-3, 5e-3, 97), np.linspace(-5e-3, 5e-3, 129), np.linspace(-5e-3, 5e-3, 65)
    rng = np.random.RandomState(5)
    X, Y, Z = np.meshgrid(x, y, z, indexing="ij")
    f = np.zeros_like(X)
    for _ in range(24):                                   # smooth random field: 24 plane waves, 1.2 - 5 mm
        k = rng.randn(3)
        k *= 2 * np.pi / (rng.uniform(1.2e-3, 5e-3) * np.linalg.norm(k))
        f += rng.randn() * np.cos(k[0] * X + k[1] * Y + k[2] * Z + rng.uniform(0, 2 * np.pi))
    ne = 1e25 * np.clip(1 + 0.3 * f / f.std(), 0, None)
    np.random.seed(12)
    s0 = orc.init_beam(4096, BEAM, 0.05e-3, 5e-3, "y")
    field = orc_c.make_field(ne, x, y, z)
    ref = orc_c.solve(field, s0, 5e-3, "y", rtol=1e-13, atol=1e-16, batch=1)[0]

    _, rf64 = _trace(tt, x, y, z, ne, s0, 5e-3, "y", "float64", 8)
    p, a, rms = _errors(rf64, ref)
    print(f"97x129x65 probing y, fp64 8 steps/cell vs C oracle: {p:.2e} m, angle {a:.1e} of rms ({rms * 1e3:.2f} mrad)")
    assert p <= 1e-5 * BEAM and a <= 1e-5
    _, rf64_2 = _trace(tt, x, y, z, ne, s0, 5e-3, "y", "float64", 2)
    _, rf32_2 = _trace(tt, x, y, z, ne, s0, 5e-3, "y", "float32", 2)
    p = np.abs(rf32_2[0::2] - rf64_2[0::2]).max()
    print(f"fp32 vs fp64 at 2 steps/cell: {p:.2e} m = {p / PIXEL_M:.1e} pixel")
    assert p <= 1e-3 * PIXEL_M
    p, a, _ = _errors(rf32_2, ref)
    print(f"fp32 2 steps/cell vs C oracle: {p:.2e} m = {p / PIXEL_M:.1e} pixel, angle {a:.1e} of rms")
    assert p <= 1e-2 * PIXEL_M                            # coarse cells (156 um = 8x the 513^3 cell): truncation


def _stretched_case():
    """the edge-case bundle of tests/test_host_kernels.py (verified there on the CPU from the kernel's source)"""
    C_LIGHT = orc.C_LIGHT
    rng = np.random.RandomState(3)
    x = np.cumsum(np.r_[0, np.geomspace(0.05e-3, 0.6e-3, 24)]) - 2e-3           # cell sizes 50 .. 600 um
    y = np.sort(np.r_[-3e-3, 3e-3, rng.uniform(-3e-3, 3e-3, 20)])
    z = np.linspace(-2e-3, 4e-3, 31) + 0.08e-3 * np.sin(np.linspace(0, 9, 31))
    X, Y, Z = np.meshgrid(x, y, z, indexing="ij")
    ne = 3e25 * (1 + 0.5 * np.sin(1500 * X) * np.cos(1100 * Y) + 0.3 * np.sin(900 * Z + 2000 * X * Y * 1e3))
    n = 64
    s0 = np.zeros((6, n))
    s0[0] = rng.uniform(x[0] + 0.5e-3, x[-1] - 0.5e-3, n)
    s0[1] = rng.uniform(-2e-3, 2e-3, n)
    s0[2] = -5e-3
    chi, phi = 2e-3 * rng.randn(n), np.pi * rng.rand(n)
    s0[3], s0[4], s0[5] = C_LIGHT * np.sin(chi) * np.cos(phi), C_LIGHT * np.sin(chi) * np.sin(phi), C_LIGHT * np.cos(chi)
    s0[0, 0], s0[1, 0] = x[5], y[7]
    s0[0, 1], s0[1, 1], s0[3, 1] = x[-1], 0.0, 0.0
    s0[2, 2] = z[4]
    s0[2, 3] = 0.5 * (z[10] + z[11])
    s0[2, 4] = z[-1]
    s0[0, 5], s0[3, 5], s0[5, 5] = x[-1] - 1e-5, 0.2 * C_LIGHT, np.sqrt(1 - 0.04) * C_LIGHT      # side exit
    s0[2, 5] = z[0]
    s0[5, 6] = -C_LIGHT                                                                           # backward
    s0[3, 7], s0[5, 7], s0[2, 7] = 0.8 * C_LIGHT, 0.6 * C_LIGHT, z[0]                            # steep
    s0[0, 8] = x[-1] + 1e-3                                                                       # misses
    return x, y, z, ne, s0, [1, 5, 6, 7, 8]


def test_rectilinear_event_marching_and_second_pass(tt):
    """tt_trace_axes variant 0 (event marching + gather second pass) against variant 2 (gather alone) and the C
    oracle on strongly stretched, asymmetric axes with rays in front of the cube, on nodes / faces, leaving
    through a side face, backward, steep, missing."""
    x, y, z, ne, s0, special = _stretched_case()
    out = {}
    for variant in (0, 2):
        cube = tt.particle_tracker.ElectronCube(x, y, z, dtype="float64", steps_per_cell=8, verbose=False)
        cube.external_ne(ne)
        cube.calc_dndr()
        assert cube._nodes is not None
        cube.kernel_variant = variant
        cube.s0 = s0
        cube.extent = 5e-3
        rf = np.asarray(cube.solve())
        out[variant] = (rf, np.asarray(cube.sf), np.asarray(cube.status), cube.ray_steps)
    (rf0, sf0, st0, n0), (rf2, sf2, st2, n2) = out[0], out[2]
    np.testing.assert_array_equal(rf0[:, special[1:]], rf2[:, special[1:]])     # the second pass IS the gather kernel
    np.testing.assert_array_equal(st0[special[1:]], st2[special[1:]])
    assert st0[8] & 8 and st0[5] & 2 and st0[7] & 16 and not np.any(st0 == 0xFF)
    marched = np.ones(s0.shape[1], bool)
    marched[special] = False
    assert np.all(st0[marched] == 1) and np.all(st2[marched] & 1)
    field = orc_c.make_field(ne, x, y, z)
    ref, sf_ref, _ = orc_c.solve(field, s0, 5e-3, "z", rtol=1e-13, atol=1e-16, batch=1, strict=False)
    p0, a0, _ = _errors(rf0[:, marched], ref[:, marched])
    p2, a2, _ = _errors(rf2[:, marched], ref[:, marched])
    print(f"stretched axes, 8 steps/cell vs C oracle: event marching {p0:.1e} m / {a0:.1e} of rms; gather {p2:.1e} m / {a2:.1e}")
    assert p0 <= 1e-9 and a0 <= 1e-6
    assert p2 <= 1e-5 * BEAM and a2 <= 1e-4
    np.testing.assert_allclose(sf0[:3, marched], sf_ref[:3, marched], rtol=0, atol=1e-8)
    assert abs(n0 - n2) <= 8 * 2                       # same plane arrivals (the general integrator may differ by a step)


def test_rectilinear_event_marching_beam_against_c_oracle(tt):
    """a full beam through a stretched 97 x 81 x 129 mesh: FP64 criterion at 2 steps/cell, float32 grid within the
    pixel bar, and the speed of the two rectilinear kernels"""
    import torch
    t = np.linspace(-1, 1, 97)
    x = 5e-3 * np.tanh(1.6 * t) / np.tanh(1.6)                       # fine in the middle, coarse outside
    y = np.linspace(-5e-3, 5e-3, 81)
    z = -5e-3 + 1e-2 * np.linspace(0, 1, 129) ** 1.5                # fine at the entry face
    X, Y, Z = np.meshgrid(x, y, z, indexing="ij")
    rng = np.random.RandomState(8)
    f = np.zeros_like(X)
    for _ in range(24):
        k = rng.randn(3)
        k *= 2 * np.pi / (rng.uniform(1.2e-3, 5e-3) * np.linalg.norm(k))
        f += rng.randn() * np.cos(k[0] * X + k[1] * Y + k[2] * Z + rng.uniform(0, 2 * np.pi))
    ne = 1e25 * np.clip(1 + 0.3 * f / f.std(), 0, None)
    np.random.seed(13)
    s0 = orc.init_beam(200_000, BEAM, 0.05e-3, 5e-3, "z")
    nref = 4096
    ref = orc_c.solve(orc_c.make_field(ne, x, y, z), s0[:, :nref], 5e-3, "z", rtol=1e-13, atol=1e-16, batch=1)[0]
    ms = {}
    for dtype, variant, spc in (("float64", 0, 2), ("float64", 2, 2), ("float32", 0, 2)):
        cube = tt.particle_tracker.ElectronCube(x, y, z, dtype=dtype, steps_per_cell=spc, verbose=False)
        cube.external_ne(ne)
        cube.calc_dndr()
        cube.kernel_variant = variant
        cube.s0 = s0
        cube.extent = 5e-3
        cube.solve()                                                  # warm-up
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rf = cube.solve()
        e1.record()
        torch.cuda.synchronize()
        ms[(dtype, variant)] = e0.elapsed_time(e1)
        assert int((cube.status.torch == 1).sum()) == s0.shape[1]
        assert cube.ray_steps == spc * 128 * s0.shape[1]
        p, a, rms = _errors(np.asarray(rf)[:, :nref], ref)
        print(f"stretched 97x81x129, {dtype} grid, variant {variant}, {spc} steps/cell: {p:.2e} m, {a:.1e} of rms "
              f"({rms * 1e3:.2f} mrad); sort + trace of 2e5 rays {ms[(dtype, variant)]:.2f} ms")
        if dtype == "float64" and variant == 0:
            assert p <= 1e-5 * BEAM and a <= 1e-5
        elif dtype == "float64":                     # gather: 2nd order where a stage straddles a u / v cell face
            assert p <= 1e-4 * BEAM and a <= 1e-3
        else:
            assert p <= 1e-3 * PIXEL_M
    print(f"event marching vs gather on the rectilinear grid: {ms[('float64', 2)] / ms[('float64', 0)]:.1f}x")


@pytest.mark.parametrize("rectilinear", [False, True])
def test_non_finite_and_resting_launch_rays_do_not_disturb_the_bundle(tt, golden, rectilinear):
    """NaN / inf positions or velocities and a ray at rest (v = 0: it never leaves, the reference integrates it to the
    time cap): the kernels must terminate, flag nothing as marched that was not, and leave every other ray of the
    bundle bit-identical.  (scipy's solve_ivp would abort the WHOLE bundle on a NaN; here the damage stays local.)"""
    g = golden("trace_grf33")
    x = g["x"] if not rectilinear else g["x"] + 2e-5 * np.sin(np.linspace(0, 7, g["x"].size))
    x = np.sort(x)
    s0 = g["s0"][:, :64].copy()
    bad = {3: (0, np.nan), 7: (5, np.nan), 11: (1, np.inf), 19: (3, -np.inf)}
    dirty = s0.copy()
    for ray, (row, val) in bad.items():
        dirty[row, ray] = val
    dirty[:, 23] = 0.0                                               # at rest in the middle of the cube
    out = {}
    for name, rays in (("clean", s0), ("dirty", dirty)):
        cube = tt.particle_tracker.ElectronCube(x, x, x, dtype="float32", steps_per_cell=2, verbose=False)
        cube.external_ne(g["ne"])
        cube.calc_dndr()
        assert (cube._nodes is not None) == rectilinear
        cube.s0 = rays
        cube.extent = float(g["extent"])
        out[name] = (np.asarray(cube.solve()), np.asarray(cube.status), np.asarray(cube.sf))
    good = np.ones(64, bool)
    good[list(bad) + [23]] = False
    np.testing.assert_array_equal(out["dirty"][0][:, good], out["clean"][0][:, good])
    np.testing.assert_array_equal(out["dirty"][1][good], out["clean"][1][good])
    np.testing.assert_array_equal(out["dirty"][2][:, good], out["clean"][2][:, good])
    st = out["dirty"][1]
    assert not np.any(st == 0xFF)                                    # nothing left in the internal "deferred" state
    for ray in bad:
        assert not np.all(np.isfinite(out["dirty"][0][:, ray])), ray  # garbage in, NaN out -- never a plausible ray
    assert st[23] & 4 and not (st[23] & 1)                           # the resting ray ran into the time cap
