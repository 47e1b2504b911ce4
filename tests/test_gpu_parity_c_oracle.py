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
