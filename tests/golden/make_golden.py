"""Generate the golden fixtures in this directory from the LIVE reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference modules are imported unmodified by path; ``matplotlib`` (absent here) is
stubbed in ``sys.modules`` because ray_transfer_matrix.py imports it at module top for its
plotting helpers.  Tight-tolerance traces go through the reference's own ``dsdt`` and
``ray_at_exit`` (``ElectronCube.solve`` hard-codes scipy's default tolerances,
particle_tracker.py:323), exactly as SURVEY section 8c describes.
"""
import io
import os
import sys
import types
import contextlib
import warnings

import numpy as np

REF = os.environ.get("TT_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path.insert(0, os.path.join(REF, "particle_tracking"))
sys.path.insert(0, os.path.join(REF, "gaussian_fields"))
warnings.simplefilter("ignore")
import particle_tracker as pt          # noqa: E402
import ray_transfer_matrix as rtm      # noqa: E402
import turboGen as tg                  # noqa: E402
from scipy.integrate import solve_ivp  # noqa: E402
import scipy                           # noqa: E402


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def tight_trace(cube, rtol, atol, batch):
    """reference dsdt + solve_ivp(rtol, atol) in small batches + reference ray_at_exit."""
    T = np.sqrt(8.0) * cube.extent / pt.c
    s0 = cube.s0
    sf = np.empty_like(s0)
    for lo in range(0, s0.shape[1], batch):
        y0 = np.ascontiguousarray(s0[:, lo:lo + batch]).flatten()
        sol = solve_ivp(lambda t, y: pt.dsdt(t, y, cube), [0, T], y0, t_eval=[0, T],
                        method="RK45", rtol=rtol, atol=atol)
        sf[:, lo:lo + batch] = sol.y[:, -1].reshape(6, -1)
    cube.sf = sf
    return cube.ray_at_exit(), sf


def save(name, **arrays):
    arrays["versions"] = np.array([np.__version__, scipy.__version__])
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def axes(n, ext=5e-3):
    return np.linspace(-ext, ext, n)


# ---------------------------------------------------------------- calc_dndr
def golden_calc_dndr():
    rng = np.random.RandomState(11)
    x, y, z = np.linspace(-3e-3, 3e-3, 12), np.linspace(-2e-3, 2e-3, 10), np.linspace(-4e-3, 4e-3, 14)
    ne = 1e27 * rng.rand(12, 10, 14)              # some voxels above ne_max*nc
    cube = pt.ElectronCube(x, y, z)
    cube.external_ne(ne.copy())
    cube.calc_dndr(lwl=1053e-9, ne_max=0.5)
    pts = np.stack([rng.uniform(-3.3e-3, 3.3e-3, 64), rng.uniform(-2.2e-3, 2.2e-3, 64),
                    rng.uniform(-4.4e-3, 4.4e-3, 64)])
    pts[:, 0] = [x[-1], y[-1], z[-1]]             # exactly on the upper faces: inside
    pts[:, 1] = [x[0], y[0], z[0]]
    pts[:, 2] = [x[3], y[4], z[5]]                # exactly on a node
    save("calc_dndr", x=x, y=y, z=z, ne=ne, ne_max=0.5, lwl=1053e-9, omega=cube.omega,
         ne_nc=cube.ne_nc, dndx=cube.dndx, dndy=cube.dndy, dndz=cube.dndz,
         pts=pts, dndr_at_pts=cube.dndr(pts))
    # uniform-linspace cube (the layout the CUDA stencil supports) with the default ne_max
    x = axes(17)
    ne = 2e26 * rng.rand(17, 17, 17)
    cube = pt.ElectronCube(x, x, x)
    cube.external_ne(ne.copy())
    cube.calc_dndr()
    save("calc_dndr_uniform", x=x, ne=ne, ne_nc=cube.ne_nc, dndx=cube.dndx, dndy=cube.dndy,
         dndz=cube.dndz)


# ---------------------------------------------------------------- init_beam
def golden_init_beam():
    out = {}
    for d in "xyz":
        x = axes(9)
        cube = pt.ElectronCube(x, x * 0.8, x * 1.2, probing_direction=d)
        np.random.seed(5)
        cube.init_beam(257, 2e-3, 5e-3)
        out["s0_" + d] = cube.s0
        out["extent_" + d] = cube.extent
    save("init_beam", **out)


# ---------------------------------------------------------------- traces
def golden_traces():
    # analytic cubes, rebuilt in the tests from (kind, params, n)
    cases = [
        ("null", dict(), 21, "z", 2e-3, 5e-3, 64),
        ("slab", dict(s=8, n_e0=1e25), 33, "z", 2e-3, 5e-3, 64),
        ("exponential_cos", dict(n_e0=2e23, Ly=1e-3, s=4e-3), 33, "z", 4e-3, 0.05e-3, 64),
        ("lens", dict(n_e0=5e25, LR=1e-3), 41, "z", 3e-3, 10e-3, 64),
        ("exponential_cos", dict(n_e0=2e23, Ly=1e-3, s=4e-3), 25, "y", 4e-3, 0.05e-3, 32),
        ("slab", dict(s=8, n_e0=1e25), 21, "x", 2e-3, 5e-3, 16),
    ]
    for i, (kind, kw, n, d, bs, div, np_) in enumerate(cases):
        x = axes(n)
        cube = pt.ElectronCube(x, x, x, probing_direction=d)
        getattr(cube, "test_" + kind)(**kw)
        cube.calc_dndr()
        np.random.seed(100 + i)
        cube.init_beam(np_, bs, div)
        rf, sf = tight_trace(cube, 1e-10, 1e-13, 32)
        save(f"trace_{i}_{kind}_{d}", kind=kind, kw_keys=np.array(list(kw.keys())),
             kw_vals=np.array(list(kw.values()), dtype=float), n=n, direction=d,
             s0=cube.s0, rf=rf, sf=sf, extent=cube.extent, rtol=1e-10, atol=1e-13)

    # Gaussian random field cube (k^-11/3), stored
    np.random.seed(7)
    f = quiet(tg.gaussian3D_FFT, 16, lambda k: k ** (-11.0 / 3.0))
    ne = 1e25 * np.clip(1 + 0.3 * f / f.std(), 0, None)
    x = axes(33)
    cube = pt.ElectronCube(x, x, x)
    cube.external_ne(ne.copy())
    cube.calc_dndr()
    np.random.seed(8)
    cube.init_beam(96, 4e-3, 0.05e-3)
    rf, sf = tight_trace(cube, 1e-10, 1e-13, 32)
    save("trace_grf33", ne=ne, x=x, s0=cube.s0, rf=rf, sf=sf, extent=cube.extent,
         rtol=1e-10, atol=1e-13)

    # the reference's own solve() (default tolerances, all rays in one call) -- the baseline path
    x = axes(25)
    cube = pt.ElectronCube(x, x, x)
    cube.test_exponential_cos(n_e0=2e23, Ly=1e-3, s=4e-3)
    cube.calc_dndr()
    np.random.seed(9)
    cube.init_beam(200, 4e-3, 0.05e-3)
    rf = quiet(cube.solve)
    save("solve_default", n=25, s0=cube.s0, rf=rf, sf=cube.sf)

    # over-critical liner: stress of the ne_max clip and strongly deflected rays
    x = axes(31)
    cube = pt.ElectronCube(x, x, x)
    cube.test_liner(n_e0=2e27, LR=1e-3)
    cube.calc_dndr()
    np.random.seed(10)
    cube.init_beam(48, 3e-3, 1e-3)
    rf, sf = tight_trace(cube, 1e-9, 1e-12, 16)
    save("trace_liner", n=31, s0=cube.s0, rf=rf, sf=sf, extent=cube.extent)


# ---------------------------------------------------------------- rectilinear (non-uniformly spaced) axes
def stretched_axes():
    """Three differently stretched axes over [-5, 5] mm: tanh clustering towards the centre, geometric growth,
    and a sinusoidal modulation of the node spacing (ratios of neighbouring cell sizes up to ~1.3)."""
    ext = 5e-3
    u = np.linspace(-1, 1, 29)
    x = ext * np.tanh(1.6 * u) / np.tanh(1.6)
    w = np.cumsum(np.concatenate([[0.0], 1.07 ** np.arange(32)]))
    y = ext * (2 * w / w[-1] - 1)
    v = np.linspace(-1, 1, 37)
    z = ext * (v + 0.12 * np.sin(2 * np.pi * v) / (2 * np.pi) * 2)
    for a in (x, y, z):
        a[0], a[-1] = -ext, ext
        assert np.all(np.diff(a) > 0)
    return x, y, z


def golden_rectilinear():
    x, y, z = stretched_axes()
    X, Y, Z = np.meshgrid(x, y, z, indexing="ij")
    rng = np.random.RandomState(21)
    ne = 2e25 * (1 + 0.5 * np.sin(2 * np.pi * X / 4e-3) * np.cos(2 * np.pi * Y / 3e-3)) * np.exp(-(Z / 3e-3) ** 2)
    ne *= 1 + 0.05 * rng.rand(*ne.shape)
    pts = np.stack([rng.uniform(-5.5e-3, 5.5e-3, 200) for _ in range(3)])
    pts[:, 0] = [x[-1], y[-1], z[-1]]
    pts[:, 1] = [x[0], y[0], z[0]]
    pts[:, 2] = [x[7], y[9], z[11]]
    out = dict(x=x, y=y, z=z, ne=ne, pts=pts)
    for d, np_, seed in (("z", 64, 31), ("y", 32, 32), ("x", 16, 33)):
        cube = pt.ElectronCube(x, y, z, probing_direction=d)
        cube.external_ne(ne.copy())
        cube.calc_dndr()
        if d == "z":
            sub = (slice(None, None, 2),) * 3            # every other node (all three axes are odd: faces included)
            out.update(dndx_sub=cube.dndx[sub], dndy_sub=cube.dndy[sub], dndz_sub=cube.dndz[sub],
                       dndr_at_pts=cube.dndr(pts))
        np.random.seed(seed)
        cube.init_beam(np_, 4e-3, 2e-3)
        rf, sf = tight_trace(cube, 1e-10, 1e-13, 32)
        out.update({f"s0_{d}": cube.s0, f"rf_{d}": rf, f"sf_{d}": sf, f"extent_{d}": cube.extent})
    save("trace_rectilinear", rtol=1e-10, atol=1e-13, **out)


# ---------------------------------------------------------------- 129^3 random cube (cube regenerated from the seed)
def golden_grf129():
    """k^-11/3 cube at 129^3 from the reference's own generator (np.random.seed(17)); 128 rays traced by the
    reference at rtol 1e-10 in batches of 32.  Only the rays are stored: the tests rebuild the cube with the
    oracle's generator from the same seed (bit-identical to the reference's, test_oracle_golden)."""
    np.random.seed(17)
    f = quiet(tg.gaussian3D_FFT, 64, lambda k: k ** (-11.0 / 3.0))
    ne = 1e25 * np.clip(1 + 0.3 * f / f.std(), 0, None)
    x = axes(129)
    cube = pt.ElectronCube(x, x, x)
    cube.external_ne(ne.copy())
    cube.calc_dndr()
    np.random.seed(18)
    cube.init_beam(128, 4e-3, 0.05e-3)
    rf, sf = tight_trace(cube, 1e-10, 1e-13, 32)
    save("trace_grf129", s0=cube.s0, rf=rf, sf=sf, extent=cube.extent, seed=17, ne_checksum=float(ne.sum()),
         rtol=1e-10, atol=1e-13)


# ---------------------------------------------------------------- spectrum diagnostic
def golden_spectrum():
    import calculate_spectrum_3d as cs
    rng = np.random.RandomState(5)
    out = {}
    np.random.seed(41)
    f = quiet(tg.gaussian3D_FFT, 12, lambda k: k ** (-11.0 / 3.0))       # 25^3
    out["f"] = f
    out["k_a"], out["s_a"] = cs.spectrum_3D_scalar(f, 1.0, k_bin_num=24)
    d = rng.randn(12, 10, 14)                                            # non-cubic, even sizes
    out["d"] = d
    out["k_b"], out["s_b"] = cs.spectrum_3D_scalar(d, 0.5, k_bin_num=16)
    save("spectrum", **out)


# ---------------------------------------------------------------- BASELINE configs[0] (C1), reduced ray count
def golden_c1():
    """100^3 test_exponential_cos cube (parameters of example_multiprocess.py:35-38), beam 4 mm,
    divergence 0.05 mrad, seed 0; 1024 rays traced by the reference at rtol 1e-10, then the
    reference's own detectors and histograms on those rays."""
    x = axes(100)
    cube = pt.ElectronCube(x, x, x)
    cube.test_exponential_cos(n_e0=2e23, Ly=1e-3, s=4e-3)
    cube.calc_dndr()
    np.random.seed(0)
    cube.init_beam(1024, 4e-3, 0.05e-3)
    rf, sf = tight_trace(cube, 1e-10, 1e-13, 32)
    out = dict(s0=cube.s0, rf=rf, sf=sf)
    dets = {"sh": (rtm.Shadowgraphy, {}), "df": (rtm.Schlieren_DF, dict(R=1)), "lf": (rtm.Schlieren_LF, dict(R=1)),
            "afr": (rtm.AFR, dict(Rs=np.arange(0, 6, .5)))}
    for k, (cls, skw) in dets.items():
        d = cls(rf)
        d.solve(**skw)
        d.histogram(bin_scale=10)
        iy, ix = np.nonzero(d.H)
        out[k + "_idx"] = np.stack([iy, ix]).astype(np.int32)
        out[k + "_cnt"] = d.H[iy, ix].astype(np.int32)
        out[k + "_rf"] = d.rf
    save("c1_expcos100", **out)


# ---------------------------------------------------------------- optics + histogram
def golden_optics():
    rng = np.random.RandomState(21)
    n = 4000
    r0 = np.zeros((4, n))
    rad = 4e-3 * np.sqrt(rng.rand(n))
    th = 2 * np.pi * rng.rand(n)
    r0[0], r0[2] = rad * np.cos(th), rad * np.sin(th)
    r0[1], r0[3] = 8e-3 * rng.randn(n), 8e-3 * rng.randn(n)     # rad; some miss the lenses / stops
    out = dict(r0=r0)
    r = rtm.m_to_mm(r0)
    out["m_to_mm"] = r
    out["lens"] = rtm.lens(r.copy(), 300.0, 150.0)
    out["sym_lens"] = rtm.sym_lens(r.copy(), 250.0)
    out["distance"] = rtm.distance(r.copy(), 123.0)
    out["circular_aperture"] = rtm.circular_aperture(r.copy(), 3.0)
    out["circular_stop"] = rtm.circular_stop(r.copy(), 3.0)
    out["annular_stop"] = rtm.annular_stop(r.copy(), 1.0, 2.5)
    out["angular_filter"] = rtm.angular_filter(r.copy(), np.arange(0, 6, 0.5))
    out["rect_aperture"] = rtm.rect_aperture(r.copy(), 2.0, 1.0)
    out["knife_edge_y_pos"] = rtm.knife_edge(r.copy(), 0.5, "y", 1)
    out["knife_edge_x_neg"] = rtm.knife_edge(r.copy(), -0.5, "x", -1)
    dets = {
        "sh": (rtm.Shadowgraphy, dict(), dict(L=400, R=25, Lx=18, Ly=13.5, focal_plane=0)),
        "sh_fp": (rtm.Shadowgraphy, dict(), dict(L=400, R=25, Lx=6, Ly=6, focal_plane=5)),
        "df": (rtm.Schlieren_DF, dict(R=3), dict(L=400, R=25, Lx=6, Ly=6)),
        "lf": (rtm.Schlieren_LF, dict(R=3), dict(L=400, R=25, Lx=6, Ly=6)),
        "afr": (rtm.AFR, dict(Rs=np.arange(0, 6, 0.5)), dict(focal_plane=5, L=100, R=25, Lx=15, Ly=10)),
    }
    for k, (cls, skw, ckw) in dets.items():
        d = cls(r0, **ckw)
        d.solve(**skw)
        d.histogram(bin_scale=25)
        out[k + "_rf"] = d.rf
        out[k + "_H"] = d.H.astype(np.int32)
        out[k + "_xedges"] = d.xedges
        out[k + "_yedges"] = d.yedges
    d = rtm.Shadowgraphy(r0)
    d.solve()
    d.histogram()                                   # default KAF-8300 binning (257, 344)
    out["sh_default_H"] = d.H.astype(np.int32)
    save("optics", **out)


# ---------------------------------------------------------------- GRF
def golden_grf():
    spec = lambda k: k ** (-11.0 / 3.0)
    out = {}
    for nd, fn, N in ((1, tg.gaussian1D_FFT, 20), (2, tg.gaussian2D_FFT, 10), (3, tg.gaussian3D_FFT, 8)):
        np.random.seed(30 + nd)
        out[f"f{nd}"] = quiet(fn, N, spec)
        out[f"N{nd}"] = N
    save("grf", **out)


if __name__ == "__main__":
    if "--c1" in sys.argv:
        golden_c1()
        sys.exit(0)
    if "--grf129" in sys.argv:
        golden_grf129()
        sys.exit(0)
    if "--rect" in sys.argv:
        golden_rectilinear()
        sys.exit(0)
    if "--spectrum" in sys.argv:
        golden_spectrum()
        sys.exit(0)
    golden_calc_dndr()
    golden_init_beam()
    golden_optics()
    golden_grf()
    golden_traces()
    golden_c1()
    golden_spectrum()
    golden_grf129()
    golden_rectilinear()
