"""Pins oracle/ref_numpy.py (the CPU restatement) to fixtures produced by the live reference
(tests/golden/make_golden.py).  CPU only."""
import glob
import os

import numpy as np
import pytest

from oracle import ref_numpy as orc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_calc_dndr_matches_reference(golden):
    g = golden("calc_dndr")
    d = orc.calc_dndr(g["ne"], g["x"], g["y"], g["z"], float(g["lwl"]), float(g["ne_max"]))
    assert d["omega"] == float(g["omega"])
    for k in ("ne_nc", "dndx", "dndy", "dndz"):
        np.testing.assert_array_equal(d[k], g[k])
    f = orc.GradientField(g["x"], g["y"], g["z"], d["dndx"], d["dndy"], d["dndz"])
    np.testing.assert_array_equal(f.dndr(g["pts"]), g["dndr_at_pts"])
    # faces are inside, beyond them the gradient is exactly zero
    assert np.any(g["dndr_at_pts"][:, 0] != 0) and np.any(g["dndr_at_pts"][:, 1] != 0)
    g = golden("calc_dndr_uniform")
    d = orc.calc_dndr(g["ne"], g["x"], g["x"], g["x"])
    for k in ("ne_nc", "dndx", "dndy", "dndz"):
        np.testing.assert_array_equal(d[k], g[k])


def test_critical_density_known_answer():
    # notebook line 427 prints nc = 1.006844605946948e+27 with c = 3e8
    omega = 2 * np.pi * 3e8 / 1053e-9
    assert orc.NC_OVER_OMEGA2 * omega**2 == pytest.approx(1.006844605946948e27, rel=1e-14)
    assert orc.critical_density()[1] == pytest.approx(1.00545e27, rel=1e-5)


def test_init_beam_matches_reference(golden):
    g = golden("init_beam")
    for d in "xyz":
        np.random.seed(5)
        s0 = orc.init_beam(257, 2e-3, 5e-3, float(g["extent_" + d]), d)
        np.testing.assert_array_equal(s0, g["s0_" + d])


def _cube_for(g):
    n = int(g["n"])
    x = np.linspace(-5e-3, 5e-3, n)
    kw = dict(zip([str(k) for k in g["kw_keys"]], [float(v) for v in g["kw_vals"]]))
    ne = orc.density(str(g["kind"]), x, x, x, **kw)
    return x, ne


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "trace_[0-9]_*.npz"))))
def test_tight_trace_matches_reference(path):
    g = np.load(path)
    x, ne = _cube_for(g)
    field = orc.make_field(ne, x, x, x)
    d = str(g["direction"])
    rf, sf, _ = orc.solve(field, g["s0"], float(g["extent"]), d, rtol=float(g["rtol"]),
                          atol=float(g["atol"]), batch=32)
    np.testing.assert_allclose(sf, g["sf"], rtol=1e-13, atol=0)
    np.testing.assert_allclose(rf, g["rf"], rtol=1e-12, atol=1e-18)


def test_null_and_slab_closed_forms(golden):
    # null: straight lines (notebook cells 3-5)
    g = golden("trace_0_null_z")
    s0, rf = g["s0"], g["rf"]
    ext = float(g["extent"])
    np.testing.assert_allclose(rf[0], s0[0] + s0[3] / s0[5] * 2 * ext, rtol=1e-12)
    np.testing.assert_allclose(rf[1], np.arctan(s0[3] / s0[5]), rtol=1e-12)
    # slab: uniform acceleration a_x = -c^2/2 * s*n_e0/(extent*nc) for the whole transit
    g = golden("trace_1_slab_z")
    s0, sf = g["s0"], g["sf"]
    nc = orc.critical_density()[1]
    a = -0.5 * orc.C_LIGHT**2 * 8 * 1e25 / (ext * nc)
    t_transit = 2 * ext / s0[5]
    np.testing.assert_allclose(sf[3], s0[3] + a * t_transit, rtol=2e-6)
    assert np.mean(g["rf"][1]) == pytest.approx(-79.3e-3, abs=0.3e-3)   # "around 80 mrad"


def test_grf_trace_and_default_solve(golden):
    g = golden("trace_grf33")
    field = orc.make_field(g["ne"], g["x"], g["x"], g["x"])
    rf, sf, _ = orc.solve(field, g["s0"][:, :32], float(g["extent"]), "z", rtol=1e-10,
                          atol=1e-13, batch=32)
    np.testing.assert_allclose(rf, g["rf"][:, :32], rtol=1e-12, atol=1e-18)
    g = golden("solve_default")
    x = np.linspace(-5e-3, 5e-3, int(g["n"]))
    ne = orc.density("exponential_cos", x, x, x, n_e0=2e23, Ly=1e-3, s=4e-3)
    rf, sf, nfev = orc.solve(orc.make_field(ne, x, x, x), g["s0"], x.max(), "z")
    np.testing.assert_allclose(rf, g["rf"], rtol=1e-12, atol=1e-18)
    assert nfev > 0


def test_liner_over_critical(golden):
    g = golden("trace_liner")
    x = np.linspace(-5e-3, 5e-3, int(g["n"]))
    ne = orc.density("liner", x, x, x, n_e0=2e27, LR=1e-3)
    rf, sf, _ = orc.solve(orc.make_field(ne, x, x, x), g["s0"][:, :16], float(g["extent"]), "z",
                          rtol=1e-9, atol=1e-12, batch=16)
    np.testing.assert_allclose(rf, g["rf"][:, :16], rtol=1e-10, atol=1e-15)


def test_optics_elements_and_detectors(golden):
    g = golden("optics")
    r0 = g["r0"]
    r = orc.m_to_mm(r0)
    np.testing.assert_array_equal(r, g["m_to_mm"])
    eq = lambda a, k: np.testing.assert_array_equal(a, g[k])
    eq(orc.lens(r.copy(), 300.0, 150.0), "lens")
    eq(orc.sym_lens(r.copy(), 250.0), "sym_lens")
    eq(orc.distance(r.copy(), 123.0), "distance")
    eq(orc.circular_aperture(r.copy(), 3.0), "circular_aperture")
    eq(orc.circular_stop(r.copy(), 3.0), "circular_stop")
    eq(orc.annular_stop(r.copy(), 1.0, 2.5), "annular_stop")
    eq(orc.angular_filter(r.copy(), np.arange(0, 6, 0.5)), "angular_filter")
    eq(orc.rect_aperture(r.copy(), 2.0, 1.0), "rect_aperture")
    eq(orc.knife_edge(r.copy(), 0.5, "y", 1), "knife_edge_y_pos")
    eq(orc.knife_edge(r.copy(), -0.5, "x", -1), "knife_edge_x_neg")
    cases = {
        "sh": ("shadowgraphy", dict(L=400, R=25, focal_plane=0), dict(Lx=18, Ly=13.5)),
        "sh_fp": ("shadowgraphy", dict(L=400, R=25, focal_plane=5), dict(Lx=6, Ly=6)),
        "df": ("schlieren_df", dict(L=400, R=25, R_stop=3), dict(Lx=6, Ly=6)),
        "lf": ("schlieren_lf", dict(L=400, R=25, R_stop=3), dict(Lx=6, Ly=6)),
        "afr": ("afr", dict(L=100, R=25, focal_plane=5, Rs=np.arange(0, 6, 0.5)), dict(Lx=15, Ly=10)),
    }
    H = {}
    for k, (kind, kw, hk) in cases.items():
        rf = orc.detector(kind, r0, **kw)
        eq(rf, k + "_rf")
        H[k], xe, ye = orc.histogram(rf, bin_scale=25, **hk)
        eq(H[k], k + "_H")
        eq(xe, k + "_xedges")
        eq(ye, k + "_yedges")
        assert H[k].shape == (2574 // 25, 3448 // 25)
    # notebook cells 16-19: dark field + light field == shadowgraphy, bin for bin
    rf_sh6 = orc.detector("shadowgraphy", r0, L=400, R=25)
    H_sh6, _, _ = orc.histogram(rf_sh6, Lx=6, Ly=6, bin_scale=25)
    np.testing.assert_array_equal(H["df"] + H["lf"], H_sh6)
    Hd, _, _ = orc.histogram(orc.detector("shadowgraphy", r0))
    eq(Hd, "sh_default_H")
    assert Hd.shape == (257, 344)
    # 4f relay is -identity at focal_plane = 0
    keep = ~np.isnan(g["sh_rf"][0])
    np.testing.assert_allclose(g["sh_rf"][0][keep], -r[0][keep], atol=1e-11)


def test_grf_matches_reference(golden):
    g = golden("grf")
    spec = lambda k: k ** (-11.0 / 3.0)
    for nd in (1, 2, 3):
        np.random.seed(30 + nd)
        f = orc.gaussian_fft(int(g[f"N{nd}"]), spec, ndim=nd)
        np.testing.assert_array_equal(f, g[f"f{nd}"])


def test_spectrum_matches_reference(golden):
    g = golden("spectrum")
    k, s = orc.spectrum_3d_scalar(g["f"], 1.0, 24)
    np.testing.assert_allclose(k, g["k_a"], rtol=1e-14)
    np.testing.assert_allclose(s, g["s_a"], rtol=1e-12, atol=1e-12 * np.nanmax(g["s_a"]), equal_nan=True)
    k, s = orc.spectrum_3d_scalar(g["d"], 0.5, 16)
    np.testing.assert_allclose(k, g["k_b"], rtol=1e-14)
    np.testing.assert_allclose(s, g["s_b"], rtol=1e-12, equal_nan=True)


def test_grf129_cube_and_first_batch(golden):
    """the 129^3 fixture: the cube regenerated from the seed is the reference's, and the oracle reproduces
    the reference's tight-tolerance trace of the first batch of 32 rays"""
    g = golden("trace_grf129")
    np.random.seed(int(g["seed"]))
    f = orc.gaussian_fft(64, lambda k: k ** (-11.0 / 3.0))
    ne = 1e25 * np.clip(1 + 0.3 * f / f.std(), 0, None)
    assert ne.sum() == float(g["ne_checksum"])
    x = np.linspace(-5e-3, 5e-3, 129)
    field = orc.make_field(ne, x, x, x)
    rf, sf, _ = orc.solve(field, g["s0"][:, :32], float(g["extent"]), "z", rtol=1e-10, atol=1e-13, batch=32)
    np.testing.assert_allclose(rf, g["rf"][:, :32], rtol=1e-12, atol=1e-18)


def test_rectilinear_axes_match_reference(golden):
    """Non-uniformly spaced axes (numpy.gradient's non-uniform stencil, bisection in the interpolator)."""
    g = golden("trace_rectilinear")
    x, y, z = g["x"], g["y"], g["z"]
    assert np.diff(x).max() / np.diff(x).min() > 3          # genuinely stretched
    d = orc.calc_dndr(g["ne"], x, y, z)
    sub = (slice(None, None, 2),) * 3
    for k in ("dndx", "dndy", "dndz"):
        np.testing.assert_array_equal(d[k][sub], g[k + "_sub"])
    f = orc.GradientField(x, y, z, d["dndx"], d["dndy"], d["dndz"])
    np.testing.assert_array_equal(f.dndr(g["pts"]), g["dndr_at_pts"])
    for dr, n in (("z", 32), ("x", 16)):
        rf, sf, _ = orc.solve(f, g["s0_" + dr][:, :n], float(g["extent_" + dr]), dr, rtol=1e-10, atol=1e-13, batch=32)
        np.testing.assert_allclose(rf, g["rf_" + dr][:, :n], rtol=1e-12, atol=1e-18)
        np.testing.assert_allclose(sf, g["sf_" + dr][:, :n], rtol=1e-13, atol=0)
