"""GPU parity at the BENCHMARKED sizes: BASELINE configs[2] (513^3, the default bench workload), configs[4]
(1025^3: a 17.2 GB float4 / 34.5 GB double4 gradient grid, 64-bit plane offsets) and configs[3] at 257^3 (packed
kernel with the passive quantities on board).  Reference: particle_tracker.py:220-256 (grid, look-up),
:312-331 (solve), :333-380 (ray_at_exit), :398-419 (dsdt), :147-210 (density set-ups), through the C oracle
(oracle/tt_oracle.c, pinned to the live reference's fixtures by tests/test_oracle_c.py) with one step sequence per
ray at rtol = 1e-13.

Tolerances (BASELINE.json north_star):
  FP64 mode : exit positions within 1e-5 of the beam radius, angles within 1e-5 of the rms angle
  FP32 mode : exit position within 1e-3 of a detector pixel (52 nm at the default bin_scale = 10)
  histogram : L1(H - H_ref) / sum(H_ref) <= 1e-3
"""
import numpy as np
import pytest

from oracle import c_oracle as orc_c
from oracle import ref_numpy as orc

pytestmark = pytest.mark.gpu

PIXEL_M = 18e-3 / (3448 // 10)
BEAM, DIV, EXTENT = 4e-3, 0.05e-3, 5e-3
SPECTRUM = lambda k: k ** (-11.0 / 3.0)      # noqa: E731


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


def _sharp(field, s0):
    """The sharp reference: scipy's RK45 (C restatement), one step sequence per ray, rtol = 1e-13, and solve_ivp's
    ``max_step`` option set to one cell's transit time.  Both departures from a plain tight-rtol run were forced by
    the 513^3 benchmark cube (ne clipped at 0 in its deepest troughs; measured on the host):
      * behind an exactly flat stretch of the field the error estimate is 0, the step grows tenfold per step, and a
        millimetre-long step whose stage points all land in flat spots again is accepted at any rtol: 21 of 8192 rays
        came back 2e-9 .. 4e-7 m off (rtol 1e-12 and 1e-8 disagreed with rtol 1e-13 on them) -- hence max_step;
      * solve_ivp integrates far beyond the exit face, where the reference's field jumps to its fill value 0
        (particle_tracker.py:238-240); to accept a step across that jump the controller needs h |jump of dv/dt| <=
        rtol |v|, which for a few rays in 10^4 is a step below 10 ulp(t): status -1, "step size underflow".  Those
        rays are integrated again at rtol = 1e-11 and, if need be, 1e-9 (still 10^4 below the FP64 criterion)."""
    ms = orc_c.cell_transit_time(field.x, field.y, field.z)
    ref = orc_c.solve(field, s0, EXTENT, "z", rtol=1e-13, atol=1e-16, batch=1, strict=False, max_step=ms)[0]
    for rtol in (1e-11, 1e-9):
        bad = np.flatnonzero(~np.isfinite(ref).all(axis=0))
        assert bad.size <= max(1, s0.shape[1] // 200), bad.size
        if bad.size:
            ref[:, bad] = orc_c.solve(field, np.ascontiguousarray(s0[:, bad]), EXTENT, "z", rtol=rtol, atol=rtol * 1e-3, batch=1,
                                      strict=False, max_step=ms)[0]
    assert np.isfinite(ref).all()
    return ref


def _bench_cube(tt, n_half):
    """the cube of bench.py: device GRF (seed 1234), ne = 1e25 clip(1 + 0.3 f / sigma, 0), float32 on the device"""
    f = tt.turboGen.gaussian3D_FFT(n_half, SPECTRUM, seed=1234, dtype="float32", return_device=True).torch
    f /= f.std()
    f.mul_(0.3).add_(1.0).clamp_(min=0).mul_(1e25)
    return f


def _cube(tt, x, ne, dtype, spc, **kw):
    cube = tt.particle_tracker.ElectronCube(x, x, x, dtype=dtype, steps_per_cell=spc, verbose=False, **kw)
    cube.external_ne(ne)
    cube.calc_dndr()
    return cube


def test_c3_bench_cube_production_kernel_against_c_oracle(tt):
    """BASELINE configs[2] = the default bench workload: 513^3 cube of the device generator, 8192 rays of the bench
    beam, the production kernel (packed FP32x2 event marching, 1 step per cell) and FP64 mode against the C oracle."""
    rtm = tt.ray_transfer_matrix
    ne_dev = _bench_cube(tt, 256)
    ne = ne_dev.cpu().numpy()                         # float32 values, identical on both sides
    M = ne.shape[0]
    assert M == 513
    x = np.linspace(-EXTENT, EXTENT, M)
    np.random.seed(21)
    s0 = orc.init_beam(8192, BEAM, DIV, EXTENT, "z")
    field = orc_c.make_field(ne.astype(np.float64), x, x, x)
    ref = _sharp(field, s0)
    del field

    cube32 = _cube(tt, x, ne_dev, "float32", 1)
    cube32.s0 = s0
    cube32.extent = EXTENT
    rf32 = np.asarray(cube32.solve())
    st = np.asarray(cube32.status)
    p, a, rms = _errors(rf32, ref)
    print(f"513^3 fp32 1 step/cell (bench setting) vs C oracle: {p:.2e} m = {p / PIXEL_M:.1e} pixel, angle {a:.1e} of rms "
          f"({rms * 1e3:.2f} mrad)")
    assert np.all(st == 1)                            # every ray marched by the event kernel to the far face
    assert cube32.ray_steps == (M - 1) * s0.shape[1]
    assert p <= 1e-3 * PIXEL_M
    sh = rtm.Shadowgraphy(cube32.rf)
    sh.solve()
    sh.histogram()
    Href, _, _ = orc.histogram(orc.detector("shadowgraphy", ref))
    l1 = np.abs(sh.H - Href).sum() / Href.sum()
    print(f"shadowgraphy of the 8192 rays: L1 distance to the oracle image {l1:.1e}")
    assert l1 <= 1e-3
    del cube32

    cube64 = _cube(tt, x, ne_dev, "float64", 4)
    cube64.s0 = s0
    cube64.extent = EXTENT
    rf64 = np.asarray(cube64.solve())
    p, a, _ = _errors(rf64, ref)
    print(f"513^3 fp64 4 steps/cell vs C oracle: {p:.2e} m = {p / BEAM:.1e} of the beam radius, angle {a:.1e} of rms")
    assert p <= 1e-5 * BEAM and a <= 1e-5


def test_c5_cube_1025_offsets_steps_and_sub_volume_oracle(tt):
    """BASELINE configs[4]: 1025^3 cube (17.2 GB float4 grid, byte offsets up to 1.7e10; 34.5 GB double4 grid).
    (1) full-beam properties: every ray marched, ray_steps = 1024 per ray, FP32 (1 step/cell) vs FP64 (2 steps/cell)
    within 1e-3 pixel, 1 vs 2 steps per cell within 1e-3 pixel;
    (2) a pencil beam through the far (x, y) corner region against the C oracle, which only needs the ne values of
    the columns the pencil visits (the reference's field is local: a trilinear look-up of central differences);
    (3) the grid at the far corner (largest offsets) against numpy.gradient on the same sub-block."""
    import torch
    ne_dev = _bench_cube(tt, 512)
    M = ne_dev.shape[0]
    assert M == 1025
    x = np.linspace(-EXTENT, EXTENT, M)
    n_rays = 1_000_000

    cube32 = _cube(tt, x, ne_dev, "float32", 1)
    assert cube32._grid.numel() * 4 > 2**34           # > 16 GiB: 32-bit element or byte offsets would wrap
    cube32.init_beam(n_rays, BEAM, DIV, seed=5)
    s0 = cube32.s0
    rf32 = cube32.solve().torch.clone()
    assert int((cube32.status.torch == 1).sum()) == n_rays
    assert cube32.ray_steps == (M - 1) * n_rays
    cube32.steps_per_cell = 2
    rf32_2 = cube32.solve().torch.clone()
    assert cube32.ray_steps == 2 * (M - 1) * n_rays
    d12 = float((rf32[0::2] - rf32_2[0::2]).abs().max())
    print(f"1025^3 fp32: 1 vs 2 steps per cell {d12:.2e} m = {d12 / PIXEL_M:.1e} pixel")
    assert d12 <= 1e-3 * PIXEL_M

    # (2) pencil beam around (+3.2 mm, +3.0 mm): columns ~ 800 .. 900 of 1025 in x and y
    np.random.seed(22)
    pen = orc.init_beam(4096, 0.3e-3, DIV, EXTENT, "z")
    pen[0] += 3.2e-3
    pen[1] += 3.0e-3
    lo, hi = 760, 960
    xs = x[lo:hi]
    sub = ne_dev[lo:hi, lo:hi, :].cpu().numpy().astype(np.float64)
    ref = _sharp(orc_c.make_field(sub, xs, xs, x), pen)
    assert np.abs(ref[0] - 3.2e-3).max() < 0.6e-3 and np.abs(ref[2] - 3.0e-3).max() < 0.6e-3   # stayed inside the sub-volume
    cube32.steps_per_cell = 1
    cube32.s0 = pen
    rfp = np.asarray(cube32.solve())
    p, a, rms = _errors(rfp, ref)
    print(f"1025^3 fp32 1 step/cell, pencil at (3.2, 3.0) mm vs C oracle: {p:.2e} m = {p / PIXEL_M:.1e} pixel, angle {a:.1e} "
          f"of rms ({rms * 1e3:.2f} mrad)")
    assert p <= 1e-3 * PIXEL_M

    # (3) gradient grid at the far corner against numpy.gradient (ElectronCube.dndr look-ups at the nodes)
    b0 = M - 24
    blk = ne_dev[b0:, b0:, b0:].cpu().numpy().astype(np.float64)
    xb = x[b0:]
    ii = np.arange(1, 23)                              # block nodes whose central stencil lies inside the block
    I, J, K = np.meshgrid(ii, ii, ii, indexing="ij")
    pts = np.stack([xb[I.ravel()], xb[J.ravel()], xb[K.ravel()]])
    got = np.asarray(cube32.dndr(pts))
    want = orc_c.make_field(blk, xb, xb, xb).dndr(pts)
    scale = np.abs(want).max()
    err = np.abs(got - want).max() / scale
    print(f"1025^3 float4 grid, far-corner block: max gradient error {err:.1e} of the largest component")
    assert err <= 2e-6                                 # float32 storage of FP64 central differences
    del cube32, rf32_2
    torch.cuda.empty_cache()

    cube64 = _cube(tt, x, ne_dev, "float64", 2)
    assert cube64._grid.numel() * 8 > 2**35
    cube64.s0 = s0
    rf64 = cube64.solve().torch
    assert int((cube64.status.torch == 1).sum()) == n_rays
    d = float((rf32[0::2] - rf64[0::2]).abs().max())
    print(f"1025^3 fp32 (1 step/cell) vs fp64 (2 steps/cell), 1e6 rays: {d:.2e} m = {d / PIXEL_M:.1e} pixel")
    assert d <= 1e-3 * PIXEL_M
    cube64.s0 = pen
    rfp64 = np.asarray(cube64.solve())
    p, a, _ = _errors(rfp64, ref)
    print(f"1025^3 fp64 2 steps/cell, pencil vs C oracle: {p:.2e} m = {p / BEAM:.1e} of the beam radius, angle {a:.1e} of rms")
    assert p <= 1e-5 * BEAM and a <= 1e-5


@pytest.mark.parametrize("face_grid", ["auto", False])
def test_c4_packed_aux_kernel_at_257(tt, face_grid):
    """BASELINE configs[3] at its size (257^3 ne + B + Te, as bench.py --workload c4 builds them): the FP32 event kernels
    with phase / Faraday rotation / absorption on board -- the face-coefficient kernel (tt_trace_faces_aux, the production
    path, face_grid="auto") and the packed corner-grid kernel (tt_trace_aux, face_grid=False) -- against the FP64 gather
    kernel, and a uniform plasma of the same size against the closed forms.  (Parity unpinned: the reference has call sites
    only.)"""
    import torch
    pt = tt.particle_tracker
    M = 257
    x = np.linspace(-EXTENT, EXTENT, M)
    ne = _bench_cube(tt, 128)
    g1 = tt.turboGen.gaussian3D_FFT(128, SPECTRUM, seed=77, dtype="float32", return_device=True).torch
    g1 /= g1.std()
    B = torch.zeros((M, M, M, 3), dtype=torch.float32, device="cuda")
    B[..., 2] = 10.0
    B[..., 0] = 2.0 * g1
    Te = 100.0 * torch.clamp(1 + 0.2 * g1, min=0.1)
    out = {}
    for dtype, spc in (("float32", 1), ("float64", 2)):
        cube = pt.ElectronCube(x, x, x, B_on=True, inv_brems=True, phaseshift=True, dtype=dtype, steps_per_cell=spc,
                               verbose=False, face_grid=face_grid)
        cube.external_ne(ne); cube.external_B(B); cube.external_Te(Te); cube.external_Z(1.0)
        cube.calc_dndr()
        cube.init_beam(200_000, BEAM, DIV, seed=8)
        rf = np.asarray(cube.solve())
        assert (cube._faces_aux is not None) == (dtype == "float32" and face_grid == "auto")
        out[dtype] = (rf, np.asarray(cube.amp), np.asarray(cube.phase), np.asarray(cube.pol), np.asarray(cube.status))
    a, b = out["float32"], out["float64"]
    assert np.all(a[4] == 1)                           # all on the packed event kernel
    dpos = np.abs(a[0][0::2] - b[0][0::2]).max()
    dph = np.abs(a[2] - b[2]).max()
    dpol = np.abs(a[3] - b[3]).max() / np.abs(b[3]).max()
    damp = np.abs(a[1] / b[1] - 1).max()
    print(f"c4 257^3: packed fp32 aux kernel vs fp64 gather kernel: pos {dpos:.1e} m = {dpos / PIXEL_M:.1e} pixel, phase {dph:.1e} "
          f"rad of {np.abs(b[2]).max():.0f}, rotation {dpol:.1e} (rel, max {np.abs(b[3]).max():.3f} rad), amplitude {damp:.1e} (rel)")
    assert dpos <= 1e-3 * PIXEL_M
    assert dph <= 2e-5 * np.abs(b[2]).max() and dpol <= 1e-4 and damp <= 1e-5

    # uniform plasma, closed forms, at the same size and through the same packed kernel
    ne_u = torch.full((M, M, M), 1e25, dtype=torch.float32, device="cuda")
    B_u = torch.zeros((M, M, M, 3), dtype=torch.float32, device="cuda")
    B_u[..., 0], B_u[..., 1], B_u[..., 2] = 0.3, -0.2, 10.0
    Te_u = torch.full((M, M, M), 100.0, dtype=torch.float32, device="cuda")
    cube = pt.ElectronCube(x, x, x, B_on=True, inv_brems=True, phaseshift=True, dtype="float32", verbose=False, face_grid=face_grid)
    cube.external_ne(ne_u); cube.external_B(B_u); cube.external_Te(Te_u); cube.external_Z(1.0)
    cube.calc_dndr()
    np.random.seed(2)
    cube.init_beam(4000, 3e-3, 5e-3)
    rf = np.asarray(cube.solve())
    assert np.all(np.asarray(cube.status) == 1)
    d = cube.s0[3:] / orc.C_LIGHT
    path = 2 * EXTENT / d[2]
    omega, nc = orc.critical_density()
    ne32, b32 = float(np.float32(1e25)), [float(np.float32(v)) for v in (0.3, -0.2, 10.0)]
    np.testing.assert_allclose(np.asarray(cube.phase), omega / orc.C_LIGHT * (np.sqrt(1 - ne32 / nc) - 1) * path, rtol=2e-6)
    Bd = b32[0] * d[0] + b32[1] * d[1] + b32[2] * d[2]
    np.testing.assert_allclose(np.asarray(cube.pol), pt.VERDET * 1053e-9**2 * ne32 * Bd * path, rtol=2e-6)
    kap = float(cube.kappa()[0, 0, 0])
    np.testing.assert_allclose(np.asarray(cube.amp), np.exp(-0.5 * kap * path), rtol=2e-6)
    np.testing.assert_allclose(rf[1], np.arctan(d[0] / d[2]), rtol=0, atol=3e-9)     # FP32 rounding of the direction components


@pytest.mark.parametrize("kind,kw", [("null", {}), ("slab", {"s": 8, "n_e0": 1e25}), ("linear_cos", {"s1": 0.3, "s2": 0.2, "n_e0": 5e24, "Ly": 2e-3}),
                                     ("liner", {"n_e0": 2e24, "LR": 2e-3}), ("lens", {"n_e0": 1e24, "LR": 1e-3}),
                                     ("exponential_cos", {"n_e0": 2e23, "Ly": 1e-3, "s": 4e-3})])
def test_density_setups_dsdt_and_solve_through_the_product_api(tt, kind, kw):
    """ElectronCube.test_* (particle_tracker.py:147-210) called on the product class: ne equal to the reference's
    formula, the module-level dsdt (:398-419) equal to the oracle's RHS on the same state, and solve() of each
    set-up within the FP64 criterion of the C oracle."""
    pt = tt.particle_tracker
    n = 101
    x = np.linspace(-EXTENT, EXTENT, n)
    cube = pt.ElectronCube(x, x, x, "z", dtype="float64", steps_per_cell=8, verbose=False)
    getattr(cube, "test_" + kind)(**kw)
    ne_ref = orc.density(kind, x, x, x, **kw)
    np.testing.assert_allclose(cube.ne, ne_ref, rtol=1e-13, atol=0)
    cube.calc_dndr()
    field = orc_c.make_field(ne_ref, x, x, x)
    np.random.seed(3)
    cube.init_beam(512, 3e-3, 2e-3)
    s0 = cube.s0.copy()
    # dsdt on a state inside the cube (launch rays pushed 3 mm in) and one partly outside
    for shift in (3e-3, 11e-3):
        s = s0.copy()
        s[2] += shift
        got = pt.dsdt(0.0, s.flatten(), cube).reshape(6, -1)
        want = np.concatenate([s[3:], field.dndr(s[:3])])
        np.testing.assert_array_equal(got[:3], want[:3])
        scale = max(np.abs(want[3:]).max(), 1e-300)
        assert np.abs(got[3:] - want[3:]).max() <= 1e-11 * scale, kind
    rf = np.asarray(cube.solve())
    ref = orc_c.solve(field, s0, EXTENT, "z", rtol=1e-13, atol=1e-16, batch=1, strict=False)[0]
    ok = np.asarray(cube.status) == 1                  # (the liner turns some rays around: those take the general kernel)
    assert ok.mean() > 0.9
    p, a, rms = _errors(rf[:, ok], ref[:, ok])
    print(f"test_{kind}: {ok.sum()} rays, {p:.1e} m, angle {a:.1e} of rms ({rms * 1e3:.3f} mrad)")
    assert p <= 1e-5 * 3e-3
    assert a <= 1e-5 or np.abs(rf[1::2, ok] - ref[1::2, ok]).max() <= 1e-9


def test_face_coefficient_kernel_against_corner_grid_kernel_and_abi(tt):
    """tt_build_face_grid / tt_trace_faces (the production path of solve() in float32 at 1 step per cell) against
    tt_trace on the same grid: 257^3 bench cube, 2e6 rays of the bench beam + a wide divergent beam whose side-exit
    rays both hand to the general kernel; sf with and without; non-cubic cells probing y."""
    import torch
    from turbulence_tracing_b200 import _lib
    pt = tt.particle_tracker
    ne = _bench_cube(tt, 128)
    M = ne.shape[0]
    x = np.linspace(-EXTENT, EXTENT, M)
    out = {}
    for fg in (True, False):
        cube = pt.ElectronCube(x, x, x, dtype="float32", verbose=False, face_grid=fg)
        cube.external_ne(ne)
        cube.calc_dndr()
        cube.init_beam(2_000_000, BEAM, DIV, seed=3)
        rf = cube.solve().torch.clone()
        assert (cube._faces is not None) == fg
        assert int((cube.status.torch == 1).sum()) == 2_000_000 and cube.ray_steps == (M - 1) * 2_000_000
        sf = cube.sf.torch.clone()
        np.random.seed(9)
        cube.s0 = orc.init_beam(200_000, 5.5e-3, 1e-2, EXTENT, "z")          # wide, divergent: side exits, misses
        rfw = cube.solve().torch.clone()
        out[fg] = (rf, sf, rfw, cube.status.torch.clone(), cube.ray_steps)
    a, b = out[True], out[False]
    d = float((a[0][0::2] - b[0][0::2]).abs().max())
    da = float((a[0][1::2] - b[0][1::2]).abs().max())
    print(f"face-coefficient kernel vs corner-grid kernel, 257^3, 2e6 rays: {d:.1e} m = {d / PIXEL_M:.1e} pixel, angles {da:.1e} rad")
    assert d <= 2e-4 * PIXEL_M and da <= 2e-7            # FP32 rounding of two evaluation orders
    # (sf: the position along the beam at time T carries the FP32 sum of 256 path-time increments)
    assert float((a[1][:3] - b[1][:3]).abs().max()) <= 1e-7 and float((a[1][3:] - b[1][3:]).abs().max()) <= 3e-6 * orc.C_LIGHT
    assert torch.equal(a[3], b[3])                      # same rays marched / handed over / missed
    assert 0 < int((a[3] != 1).sum()) < 200_000 and a[4] == b[4] or abs(a[4] - b[4]) <= 1e-4 * b[4]
    m = torch.isfinite(b[2]).all(dim=0)
    assert torch.equal(m, torch.isfinite(a[2]).all(dim=0))
    dw = float((a[2][0::2, m] - b[2][0::2, m]).abs().max())
    print(f"wide beam (side exits through the general kernel): {dw:.1e} m")
    assert dw <= 2e-4 * PIXEL_M

    # non-cubic cells, probing y, against the C oracle through the product API
    xx, yy, zz = np.linspace(-5e-3, 5e-3, 97), np.linspace(-5e-3, 5e-3, 129), np.linspace(-5e-3, 5e-3, 81)
    ne2 = orc.density("exponential_cos", xx, yy, zz, n_e0=3e24, Ly=2e-3, s=4e-3)
    cube = pt.ElectronCube(xx, yy, zz, "y", dtype="float32", verbose=False, face_grid=True)
    cube.external_ne(ne2)
    cube.calc_dndr()
    np.random.seed(6)
    cube.init_beam(4096, 3e-3, 1e-3)
    s0 = cube.s0.copy()
    rf = np.asarray(cube.solve())
    assert np.all(np.asarray(cube.status) == 1)
    ref = orc_c.solve(orc_c.make_field(ne2, xx, yy, zz), s0, EXTENT, "y", rtol=1e-13, atol=1e-16, batch=1, strict=False)[0]
    ok = np.isfinite(ref).all(axis=0)
    p, ang, _ = _errors(rf[:, ok], ref[:, ok])
    print(f"face kernel, non-cubic cells, probing y vs C oracle: {p:.2e} m = {p / PIXEL_M:.1e} pixel")
    assert ok.mean() > 0.99 and p <= 1e-3 * PIXEL_M

    # face_grid=True where it does not apply is an error, not a silent switch
    c64 = pt.ElectronCube(x, x, x, dtype="float64", verbose=False, face_grid=True)
    c64.external_ne(ne)
    c64.calc_dndr()
    c64.init_beam(10, BEAM, DIV, seed=1)
    with pytest.raises(ValueError):
        c64.solve()


@pytest.mark.parametrize("bin_scale,n", [(10, 3_000_000), (10, 2_999_999), (10, 70_001), (5, 3_000_000), (1, 3_000_000)])
def test_privatised_histogram_equals_numpy_and_the_general_kernel(tt, bin_scale, n):
    """optics_hist_smem16_kernel (the whole image as 16-bit counters in one CTA's shared memory; taken for >= 65536 rays in
    storage order when the image fits: bin_scale 10; at 5 and 1 it does not fit -> general kernel) against
    numpy.histogram2d on the same detector-plane rays and against the general kernel (forced by an identity permutation):
    3e6 unsorted rays with NaNs (in x only, in theta only), rays on bin edges and outside the detector; an odd ray count
    takes the kernel's scalar-load instantiation, 70 001 rays leave most CTAs a partial or empty batch."""
    import torch
    rtm = tt.ray_transfer_matrix
    rng = np.random.default_rng(12)
    r0 = np.empty((4, n))
    r0[0] = rng.uniform(-11e-3, 11e-3, n)          # m; the detector is 18 x 13.5 mm: some rays miss it
    r0[2] = rng.uniform(-8e-3, 8e-3, n)
    r0[1] = rng.normal(0, 2e-3, n)
    r0[3] = rng.normal(0, 2e-3, n)
    r0[0, :1000] = np.nan                          # dropped like numpy does after the reference's NaN filter
    r0[1, 2000:2500] = np.nan                      # a NaN angle makes the whole column NaN at the first element
    r0[0, 1000:1010] = 9e-3                        # last edge: inclusive
    r0[0, 1010:1020] = -9e-3
    sh = rtm.Shadowgraphy(r0)
    sh.solve()
    sh.histogram(bin_scale=bin_scale)
    H = sh.H.copy()
    rf = np.asarray(sh.rf)
    ok = ~np.isnan(rf[0]) & ~np.isnan(rf[2])
    Href = np.histogram2d(rf[0][ok], rf[2][ok], bins=[3448 // bin_scale, 2574 // bin_scale],
                          range=[[-9.0, 9.0], [-6.75, 6.75]])[0].T
    np.testing.assert_array_equal(H, Href)
    sh2 = rtm.Shadowgraphy(r0)
    sh2.use_ray_order(torch.arange(n, dtype=torch.int32, device="cuda"))      # identity order: the general kernel
    sh2.solve()
    sh2.histogram(bin_scale=bin_scale)
    np.testing.assert_array_equal(sh2.H, H)
    assert H.sum() > 0.6 * n


def test_privatised_histogram_counter_overflow(tt):
    """A focused beam: 12e6 of 16e6 rays fall into TWO neighbouring bins that share one 32-bit word of the privatised image
    (and 2e6 into the last bin of the image), > 32767 per CTA and bin, so every CTA moves 0x8000-blocks of counts to the
    global image many times while its neighbours keep adding to the same word; the image must still equal
    numpy.histogram2d's, count for count."""
    rtm = tt.ray_transfer_matrix
    rng = np.random.default_rng(5)
    n = 16_000_000
    r0 = np.zeros((4, n))
    r0[0] = rng.uniform(-8.9e-3, 8.9e-3, n)
    r0[2] = rng.uniform(-6.7e-3, 6.7e-3, n)
    dx = 18e-3 / 344                                # one bin in x at bin_scale = 10 (m)
    r0[0, :6_000_000] = 0.25 * dx;  r0[2, :6_000_000] = 1e-5       # bin (172, 128): even ...
    r0[0, 6_000_000:12_000_000] = 1.25 * dx; r0[2, 6_000_000:12_000_000] = 1e-5      # ... and its odd neighbour
    r0[0, 12_000_000:14_000_000] = 8.99e-3; r0[2, 12_000_000:14_000_000] = 6.74e-3    # last bin of the image
    perm = rng.permutation(n)
    r0 = np.ascontiguousarray(r0[:, perm])
    sh = rtm.Shadowgraphy(r0)
    sh.solve()
    sh.histogram(bin_scale=10)
    rf = np.asarray(sh.rf)
    Href = np.histogram2d(rf[0], rf[2], bins=[344, 257], range=[[-9.0, 9.0], [-6.75, 6.75]])[0].T
    np.testing.assert_array_equal(sh.H, Href)
    assert sh.H.sum() == n and sh.H.max() >= 6_000_000


@pytest.mark.parametrize("direction", ["x", "y", "z"])
@pytest.mark.parametrize("dt", ["float64", "float32"])
def test_fused_aux_grid_equals_the_elementwise_composition(tt, direction, dt):
    """tt_build_aux_grid ((B_u, B_v, B_w, kappa) per node in the gradient grid's layout, one pass) against the composition
    it replaced: kappa() (the documented formula, element-wise tensor operations) and the frame permutation of B, on a
    non-cubic cube with Te and Z as cubes, as scalars, with a fixed Coulomb logarithm, with B only and with kappa only."""
    import torch
    pt = tt.particle_tracker
    rng = np.random.default_rng(8)
    x, y, z = np.linspace(-5e-3, 5e-3, 23), np.linspace(-4e-3, 4e-3, 31), np.linspace(-5e-3, 5e-3, 18)
    shape = (23, 31, 18)
    ne = (1e25 * rng.uniform(0.0, 1.2, shape)).astype(dt)                 # some nodes above 0 density only; clip is at ne/nc
    B = rng.standard_normal(shape + (3,)).astype(dt) * 5
    Te = rng.uniform(20, 300, shape).astype(dt)
    Z = rng.uniform(1, 6, shape).astype(dt)
    fa = {"z": (0, 1, 2), "y": (0, 2, 1), "x": (1, 2, 0)}[direction]
    for case in ("cubes", "scalars", "lnL", "B only", "kappa only"):
        cube = pt.ElectronCube(x, y, z, direction, B_on=case != "kappa only", inv_brems=case != "B only", phaseshift=True,
                               dtype=dt, verbose=False)
        cube.external_ne(ne)
        if case != "kappa only":
            cube.external_B(B)
        if case != "B only":
            cube.external_Te(Te if case != "scalars" else 120.0)
            cube.external_Z(Z if case != "scalars" else 3.0)
        if case == "lnL":
            cube.coulomb_log = 7.5
        cube.calc_dndr()
        a = cube._aux_grid()
        assert tuple(a.shape) == (shape[fa[2]], shape[fa[1]], shape[fa[0]], 4) and str(a.dtype) == "torch." + dt
        got = a.double().cpu().numpy()
        if case == "kappa only":
            assert not got[..., :3].any()
        else:
            want_B = np.stack([B[..., fa[0]], B[..., fa[1]], B[..., fa[2]]], axis=-1).astype(np.float64)
            np.testing.assert_array_equal(got[..., :3], want_B.transpose(fa[2], fa[1], fa[0], 3))
        if case == "B only":
            assert not got[..., 3].any()
        else:
            kap = cube.kappa().double().cpu().numpy().transpose(fa[2], fa[1], fa[0])
            assert np.isfinite(kap).all() and kap.max() > 0
            np.testing.assert_allclose(got[..., 3], kap, rtol=2e-7 if dt == "float32" else 1e-14, atol=0)


def test_pageable_upload_through_the_staging_ring(tt, monkeypatch):
    """tt_h2d_pageable (worker threads -> pinned ring -> one DMA per piece) delivers the bytes of a pageable array: sizes that
    are not a multiple of the 4 MB piece, more pieces than ring slots (the slots are reused behind their DMA events), 1 / 3 /
    default worker threads, two calls back to back on different streams; small and pinned sources take the plain copy."""
    import torch
    lib = tt._lib
    rng = np.random.default_rng(2)
    a = rng.integers(0, 2**62, size=(300 * (1 << 20) + 12345) // 8, dtype=np.int64)       # 300 MB + a ragged tail: 76 pieces > 32 slots
    src = torch.from_numpy(a)
    assert not src.is_pinned()
    for threads in ("1", "3", None):
        if threads is None:
            monkeypatch.delenv("TT_H2D_THREADS", raising=False)
        else:
            monkeypatch.setenv("TT_H2D_THREADS", threads)
        dst = torch.zeros(a.shape, dtype=torch.int64, device="cuda")
        lib.h2d(dst, src)
        torch.cuda.synchronize()
        assert torch.equal(dst.cpu(), src), threads
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    d1 = torch.zeros_like(dst)
    d2 = torch.zeros(40 * (1 << 20) // 8, dtype=torch.int64, device="cuda")
    with torch.cuda.stream(s1):
        lib.h2d(d1, src)
    with torch.cuda.stream(s2):
        lib.h2d(d2, src[: d2.numel()])
    torch.cuda.synchronize()
    assert torch.equal(d1.cpu(), src) and torch.equal(d2.cpu(), src[: d2.numel()])
    small = torch.arange(1000, dtype=torch.float64)
    assert torch.equal(lib.h2d(None, small).cpu(), small)
    pinned = src[:5_000_000].clone().pin_memory()
    assert torch.equal(lib.h2d(None, pinned).cpu(), pinned)
    # the solve() pipeline with plain numpy rays takes this path: the result equals that of device-resident rays
    pt = tt.particle_tracker
    x = np.linspace(-5e-3, 5e-3, 33)
    cube = pt.ElectronCube(x, x, x, keep_sf=False, verbose=False)
    cube.external_ne(1e25 * (1 + 0.2 * np.sin(600 * x)[:, None, None] * np.cos(500 * x)[None, :, None]) * np.ones((33, 33, 33)))
    cube.calc_dndr()
    cube.init_beam(6_000_000, 4e-3, 0.05e-3, seed=3)
    s0_dev = cube.s0
    rf_dev = np.asarray(cube.solve()).copy()
    cube.s0 = np.asarray(s0_dev).copy()               # pageable numpy, 288 MB: rows of 48 MB each
    cube.pipeline_first_rays = 100_000
    rf_host = np.asarray(cube.solve())
    np.testing.assert_array_equal(rf_host, rf_dev)
