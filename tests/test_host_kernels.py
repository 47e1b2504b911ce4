"""CPU run of GPU kernel source.  The per-ray body of the rectilinear event-marching kernel
(csrc/trace_axes_event.cuh) is written for host AND device; tests/host/axes_event_host.cu wraps it in a loop over
rays and this test compiles that with nvcc for the host and checks it against the C oracle and the live
reference's fixture -- the same source the B200 runs, verified without a GPU.  (The GPU suite repeats the checks
through the C ABI.)"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import c_oracle as orc_c

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
C_LIGHT = 299792458.0
DEFERRED, EXIT_FACE = 0xFF, 1


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    out = str(tmp_path_factory.mktemp("host") / "axes_event_host.so")
    subprocess.run([nvcc, "-std=c++17", "-O1", "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a",
                    "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "turbulence_tracing_b200", "csrc"),
                    os.path.join(ROOT, "tests", "host", "axes_event_host.cu"), "-o", out], check=True)
    lib = C.CDLL(out)
    vp = C.c_void_p
    lib.host_axes_event.argtypes = [vp, C.c_int, C.POINTER(C.c_int * 3), vp, vp, vp, C.c_int, C.c_double, C.c_double, C.c_int,
                                    vp, C.c_long, vp, vp, vp, C.POINTER(C.c_ulonglong), C.POINTER(C.c_long)]
    lib.host_axes_event.restype = C.c_int
    return lib


FRAME = {2: (0, 1, 2), 1: (0, 2, 1), 0: (1, 2, 0)}          # (u, v, w) -> xyz, csrc/common.cuh frame_of


def _grid4(ne, x, y, z, par, dtype):
    """the interleaved grid of tt_calc_dndr_axes: [iw][iv][iu] of (g_u, g_v, g_w, ne/nc), g = dnd?/c^2"""
    d = orc_c.calc_dndr(ne, x, y, z)
    fa = FRAME[par]
    comp = [d["dndx"], d["dndy"], d["dndz"]]
    G = np.empty(tuple(ne.shape[a] for a in (fa[2], fa[1], fa[0])) + (4,), dtype=dtype)
    for k in range(3):
        G[..., k] = (comp[fa[k]] / C_LIGHT**2).transpose(fa[2], fa[1], fa[0])
    G[..., 3] = d["ne_nc"].transpose(fa[2], fa[1], fa[0])
    return np.ascontiguousarray(G)


def _run(lib, G, x, y, z, par, extent, s0, spc, want_sf=True):
    n = s0.shape[1]
    s0 = np.ascontiguousarray(s0, dtype=np.float64)
    rf, sf = np.full((4, n), np.nan), np.full((6, n), np.nan)
    status = np.zeros(n, dtype=np.uint8)
    steps, nd = C.c_ulonglong(), C.c_long()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    x, y, z = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z))
    rc = lib.host_axes_event(p(G), 0 if G.dtype == np.float32 else 1, C.byref((C.c_int * 3)(x.size, y.size, z.size)),
                             p(x), p(y), p(z), par, float(extent), float(np.sqrt(8.0) * extent), spc, p(s0), n, p(rf),
                             p(sf) if want_sf else None, p(status), C.byref(steps), C.byref(nd))
    assert rc == 0
    return rf, sf, status, steps.value, nd.value


def _errors(rf, ref):
    rms = max(np.sqrt(np.mean(ref[1] ** 2 + ref[3] ** 2)), 1e-6)
    return np.abs(rf[0::2] - ref[0::2]).max(), np.abs(rf[1::2] - ref[1::2]).max() / rms


def test_rectilinear_event_body_matches_reference_fixture_and_c_oracle(host_lib, golden):
    g = golden("trace_rectilinear")                 # 29 x 33 x 37 nodes, tanh / geometric / sinusoidal stretching
    x, y, z, ne = g["x"], g["y"], g["z"], g["ne"]
    field = orc_c.make_field(ne, x, y, z)
    for dr, par in (("z", 2), ("y", 1)):
        s0, ext = g["s0_" + dr], float(g["extent_" + dr])
        ref = orc_c.solve(field, s0, ext, dr, rtol=1e-13, atol=1e-16, batch=1)
        errs = {}
        for spc in (1, 2, 4, 8):
            rf, sf, status, steps, nd = _run(host_lib, _grid4(ne, x, y, z, par, np.float64), x, y, z, par, ext, s0, spc)
            assert nd == 0 and np.all(status == EXIT_FACE)
            assert steps == spc * (g["xyz"[par]].size - 1) * s0.shape[1]
            errs[spc] = _errors(rf, ref[0])
        print(dr, errs)
        assert errs[8][0] <= 1e-5 * 4e-3 and errs[8][1] <= 1e-5          # the FP64 criterion, with margin below
        assert errs[8][0] <= 1e-9 and errs[8][1] <= 1e-6
        assert errs[2][1] < errs[1][1] / 6                               # 4th order: no kink is straddled
        # against the live reference's rays (own error of the rtol = 1e-10 fixture: ~1e-6 of the rms angle)
        p, a = _errors(rf, g["rf_" + dr])
        assert p <= 1e-5 * 4e-3 and a <= 1e-5
        # state at time T, as the reference stores it
        np.testing.assert_allclose(sf[:3], ref[1][:3], rtol=0, atol=1e-8)
        np.testing.assert_allclose(sf[3:], ref[1][3:], rtol=0, atol=1e-6 * C_LIGHT)
        # float32 grid, FP64 arithmetic
        rf32 = _run(host_lib, _grid4(ne, x, y, z, par, np.float32), x, y, z, par, ext, s0, 8)[0]
        assert _errors(rf32, ref[0])[0] <= 1e-3 * 52e-6
    # probing 'x': the reference launches ON the far face (+extent) moving away from the cube (quirk kept, SURVEY
    # section 7.9): nothing to march, the ray leaves at once and undeflected
    rf, sf, status, steps, nd = _run(host_lib, _grid4(ne, x, y, z, 0, np.float64), x, y, z, 0, float(g["extent_x"]),
                                     g["s0_x"], 2)
    assert nd == 0 and steps == 0 and np.all(status == EXIT_FACE)
    p, a = _errors(rf, g["rf_x"])
    assert p <= 1e-5 * 4e-3 and a <= 1e-5


def test_rectilinear_event_body_edge_cases(host_lib):
    """asymmetric, strongly stretched axes; rays launched in front of the cube, on nodes and faces, towards a side
    face, backwards, steep: marched rays agree with the C oracle, everything else is handed to the second pass"""
    rng = np.random.RandomState(3)
    x = np.cumsum(np.r_[0, np.geomspace(0.05e-3, 0.6e-3, 24)]) - 2e-3           # cell sizes 50 .. 600 um
    y = np.sort(np.r_[-3e-3, 3e-3, rng.uniform(-3e-3, 3e-3, 20)])
    z = np.linspace(-2e-3, 4e-3, 31) + 0.08e-3 * np.sin(np.linspace(0, 9, 31))
    X, Y, Z = np.meshgrid(x, y, z, indexing="ij")
    ne = 3e25 * (1 + 0.5 * np.sin(1500 * X) * np.cos(1100 * Y) + 0.3 * np.sin(900 * Z + 2000 * X * Y * 1e3))
    field = orc_c.make_field(ne, x, y, z)
    ext = 5e-3                                     # launch plane at -extent: 3 mm in front of the cube
    n = 64
    s0 = np.zeros((6, n))
    s0[0] = rng.uniform(x[0] + 0.5e-3, x[-1] - 0.5e-3, n)
    s0[1] = rng.uniform(-2e-3, 2e-3, n)
    s0[2] = -ext
    chi, phi = 2e-3 * rng.randn(n), np.pi * rng.rand(n)
    s0[3], s0[4], s0[5] = C_LIGHT * np.sin(chi) * np.cos(phi), C_LIGHT * np.sin(chi) * np.sin(phi), C_LIGHT * np.cos(chi)
    s0[0, 0], s0[1, 0] = x[5], y[7]                # exactly on a node column
    s0[0, 1], s0[1, 1] = x[-1], 0.0                # on the upper x face, flying along it
    s0[3, 1] = 0.0
    s0[2, 2] = z[4]                                # launched inside the cube, on a node plane
    s0[2, 3] = 0.5 * (z[10] + z[11])               # inside, between planes
    s0[2, 4] = z[-1]                               # already on the far face
    special = {5: "side", 6: "backward", 7: "steep", 8: "outside"}
    s0[0, 5], s0[3, 5], s0[5, 5] = x[-1] - 1e-5, 0.2 * C_LIGHT, np.sqrt(1 - 0.04) * C_LIGHT      # leaves through +x
    s0[5, 6] = -C_LIGHT                            # flying away
    s0[3, 7], s0[5, 7] = 0.8 * C_LIGHT, 0.6 * C_LIGHT                                            # d_w < 0.75
    s0[0, 8] = x[-1] + 1e-3                        # misses the cube
    ref = orc_c.solve(field, s0, ext, "z", rtol=1e-13, atol=1e-16, batch=1, strict=False)[0]
    G = _grid4(ne, x, y, z, 2, np.float64)
    rf, sf, status, steps, nd = _run(host_lib, G, x, y, z, 2, ext, s0, 8)
    marched = status == EXIT_FACE
    # ray 1 starts ON the side face: it is marched if the field pushes it inwards, handed over if it leaves
    handed = set(np.flatnonzero(~marched).tolist())
    assert set(special) <= handed <= set(special) | {1}, (handed, status[~marched])
    assert np.all(status[~marched] == DEFERRED) and nd == len(handed)
    assert np.all(np.isnan(rf[:, ~marched]))       # untouched: the second pass writes them
    p, a = _errors(rf[:, marched], ref[:, marched])
    print(f"edge cases: {p:.2e} m, {a:.2e} of the rms angle")
    assert p <= 1e-9 and a <= 1e-6
    assert rf[0, 4] == pytest.approx(ref[0, 4], abs=1e-15) and steps > 0
    # without an sf buffer
    rf2 = _run(host_lib, G, x, y, z, 2, ext, s0, 8, want_sf=False)[0]
    np.testing.assert_array_equal(rf2[:, marched], rf[:, marched])
