"""CPU-only tests: the C-ABI library loads and exports every symbol declared in include/tt_b200.h,
the product never routes through the oracle, host-side logic (axis checks, element programs,
spectrum table, sharding) behaves, and the compute entry points fail loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tt_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tt_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from turbulence_tracing_b200 import _lib
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from turbulence_tracing_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in tt_b200.h but not exported"
    assert set(syms) == set(_lib.PROTOTYPES), "ctypes prototypes out of sync with the header"
    assert lib.tt_abi_version() == 1
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for s in syms:
        assert re.search(rf"\bT {s}\b", out), f"{s} not a defined text symbol"


def test_library_is_sm100a_only():
    from turbulence_tracing_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_struct_layouts_match_header(lib):
    from turbulence_tracing_b200 import _lib
    # tt_optic {int, int, double, double}; tt_trace_params as declared
    assert C.sizeof(_lib.Optic) == 24
    assert _lib.TraceParams.origin_xyz.offset == 16 and _lib.TraceParams.par.offset == 64
    assert _lib.TraceParams.extent.offset == 72 and _lib.TraceParams.dtype.offset == 92
    assert C.sizeof(_lib.TraceParams) == 104


def test_argument_errors_are_codes_not_exceptions(lib):
    from turbulence_tracing_b200 import _lib
    n = _lib.i3((4, 4, 4))
    h = _lib.d3((1.0, 1.0, 1.0))
    assert lib.tt_calc_dndr(None, 0, n, h, 2, 1e27, 1.0, None, 0, None) == 1
    assert b"null" in lib.tt_last_error()
    assert lib.tt_calc_dndr(C.c_void_p(8), 0, n, h, 5, 1e27, 1.0, C.c_void_p(8), 0, None) == 1
    assert b"par" in lib.tt_last_error()
    need = C.c_size_t(0)
    assert lib.tt_sort_rays_workspace(-1, C.byref(need)) == 1
    prog = (_lib.Optic * 1)()
    prog[0].op = 99
    assert lib.tt_optics_hist(C.c_void_p(8), 10, 1.0, prog, 1, None, 0, None, 0, None, C.c_void_p(8), None) == 1
    assert b"unknown op" in lib.tt_last_error()
    # round 2: staged upload and aux-grid builder
    assert lib.tt_h2d_pageable(None, None, 0, None) == 0                      # nothing to copy
    assert lib.tt_h2d_pageable(None, C.c_void_p(8), 16, None) == 1 and b"null" in lib.tt_last_error()
    nan = float("nan")
    assert lib.tt_build_aux_grid(C.c_void_p(8), 0, None, 100.0, None, 1.0, None, 0, n, 2, 1e27, 1.0, 1e15, nan, 0,
                                 C.c_void_p(8), 0, None) == 1 and b"nothing to build" in lib.tt_last_error()
    assert lib.tt_build_aux_grid(C.c_void_p(8), 0, None, 0.0, None, 1.0, None, 0, n, 2, 1e27, 1.0, 1e15, nan, 1,
                                 C.c_void_p(8), 0, None) == 1 and b"Te" in lib.tt_last_error()
    assert lib.tt_build_aux_grid(C.c_void_p(8), 7, None, 100.0, None, 1.0, None, 0, n, 2, 1e27, 1.0, 1e15, nan, 1,
                                 C.c_void_p(8), 0, None) == 1 and b"dtype" in lib.tt_last_error()


def test_no_gpu_means_loud_failure_not_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for the CPU-only container")
    from turbulence_tracing_b200 import particle_tracker as pt, ray_transfer_matrix as rtm, turboGen as tg
    from turbulence_tracing_b200 import TTError
    x = np.linspace(-1e-3, 1e-3, 5)
    cube = pt.ElectronCube(x, x, x)
    cube.external_ne(np.zeros((5, 5, 5)))
    with pytest.raises(TTError):
        cube.calc_dndr()
    with pytest.raises(TTError):
        cube.test_slab()
    with pytest.raises(TTError):
        rtm.Shadowgraphy(np.zeros((4, 3)))
    with pytest.raises(TTError):
        tg.gaussian3D_FFT(2, lambda k: k**-2.0)
    assert lib.tt_device_count() <= 0
    ne = np.zeros((5, 5, 5))
    s0 = np.zeros((6, 1))
    rf = np.zeros((4, 1))
    from turbulence_tracing_b200 import _lib
    rc = lib.tt_solve_host(ne.ctypes.data, _lib.i3((5, 5, 5)), _lib.d3((0, 0, 0)), _lib.d3((1, 1, 1)), 2, 1e27, 1.0,
                           1.0, 1, 0, s0.ctypes.data, 1, rf.ctypes.data, None, None)
    assert rc == 2 and b"no CPU fallback" in lib.tt_last_error() or b"CUDA" in lib.tt_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "turbulence_tracing_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "ref_numpy" not in src, f
    code = "import sys; import turbulence_tracing_b200; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)


def test_host_init_beam_matches_reference_draws(golden):
    from turbulence_tracing_b200 import particle_tracker as pt
    g = golden("init_beam")
    x = np.linspace(-5e-3, 5e-3, 9)
    for d in "xyz":
        cube = pt.ElectronCube(x, x * 0.8, x * 1.2, probing_direction=d)
        np.random.seed(5)
        cube.init_beam(257, 2e-3, 5e-3)
        np.testing.assert_array_equal(cube.s0, g["s0_" + d])
        assert cube.extent == float(g["extent_" + d])
    np.random.seed(5)
    s0 = pt.init_beam(257, 2e-3, 5e-3, 6e-3, "z")           # stale module-level API of the examples
    np.testing.assert_array_equal(s0, g["s0_z"])


def test_axis_validation():
    from turbulence_tracing_b200 import particle_tracker as pt
    x = np.linspace(-1e-3, 1e-3, 9)
    o, h = pt._uniform_spacing(x, "x")
    assert o == x[0] and h == pytest.approx(2.5e-4)
    o, h = pt._uniform_spacing(np.arange(-5e-3, 5e-3, 50e-6), "x")     # notebook's simulation axes
    assert h == pytest.approx(50e-6)
    bad = x.copy()
    bad[4] += 1e-6
    with pytest.raises(NotImplementedError):
        pt._uniform_spacing(bad, "x")
    assert pt._axis_spacing(bad, "x")[2] is False and pt._axis_spacing(x, "x")[2] is True
    assert pt.ElectronCube(x, bad, x)._geometry()[2] is True       # one stretched axis -> rectilinear kernels
    assert pt.ElectronCube(x, x, x)._geometry()[2] is False
    with pytest.raises(ValueError):
        pt._uniform_spacing(x[::-1], "x")
    dup = x.copy()
    dup[3] = dup[2]
    with pytest.raises(ValueError):
        pt._axis_spacing(dup, "x")
    cube = pt.ElectronCube(x, x, x, probing_direction="q")
    with pytest.raises(ValueError):
        _ = cube._par
    # both call styles: current reference (direction 4th) and the examples' older one (extent 4th)
    assert pt.ElectronCube(x, x, x, "y").probing_direction == "y"
    assert pt.ElectronCube(x, x, x).probing_direction == "z"
    c = pt.ElectronCube(x, x, x, 1e-3, B_on=True, inv_brems=False, phaseshift=True, probing_direction="x")
    assert c.probing_direction == "x" and c.B_on and c.phaseshift and c.extent_x == 1e-3
    with pytest.raises(TypeError):
        pt.ElectronCube(x, x, x, "y", probing_direction="z")


def test_detector_programs_follow_the_reference_sequences():
    from turbulence_tracing_b200 import ray_transfer_matrix as rtm, _lib
    D, Lz, AP, ST, AN = _lib.OP_DISTANCE, _lib.OP_LENS, _lib.OP_CIRC_APERTURE, _lib.OP_CIRC_STOP, _lib.OP_ANNULAR_STOP

    class Probe:
        focal_plane, L, R = 5, 400, 25
        def _set_program(self, p):
            self.p = p
    p = Probe(); rtm.Shadowgraphy.solve(p)
    assert [o[0] for o in p.p] == [D, AP, Lz, D, AP, Lz, D]
    assert [o[1] for o in p.p] == [395, 25, 400, 800, 25, 400, 400]
    p = Probe(); rtm.Schlieren_DF.solve(p, R=3)
    assert [o[0] for o in p.p] == [D, AP, Lz, D, ST, D, AP, Lz, D] and p.p[4][1] == 3
    p = Probe(); rtm.Schlieren_LF.solve(p, R=2)
    assert [o[0] for o in p.p] == [D, AP, Lz, D, AP, D, AP, Lz, D] and p.p[4][1] == 2
    p = Probe(); rtm.AFR.solve(p, np.arange(0, 6, 0.5))
    assert [o[0] for o in p.p] == [D, AP, Lz, D] + [AN] * 6 + [D, AP, Lz, D]
    assert p.p[0][1] == 195 and p.p[2][1:] == (200, 200) and p.p[3][1] == 100
    assert p.p[4][1:] == (0.0, 0.5) and p.p[9][1:] == (5.0, 5.5)
    with pytest.raises(ValueError):
        rtm._op_knife(0.0, "x", 0)


def test_spectrum_table():
    from turbulence_tracing_b200 import turboGen as tg
    N = 4
    lut = tg._sqrt_spectrum_table(N, lambda k: k ** (-11.0 / 3.0))
    assert lut.shape == (3 * N * N + 1,) and lut[0] == 0
    q = np.arange(1, 3 * N * N + 1)
    np.testing.assert_allclose(lut[1:], (np.sqrt(q) / (2 * N + 1)) ** (-11.0 / 6.0), rtol=1e-14)
    assert np.all(tg._sqrt_spectrum_table(N, lambda k: 2.0)[1:] == np.sqrt(2.0))   # scalar-valued k_func


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (CPU arm): one JSON line with the contract's keys."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                          "--steps", "1", "--warmup", "0", "--cpu-rays", "40"], capture_output=True, text=True,
                         timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "impl", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "ray-steps/s" and line["value"] > 0
    # the reference's own modules (oracle/_ref after `make -C oracle ref`, or /root/reference in the build container);
    # "port" only where neither exists
    from oracle import reference_live as live
    assert line["cpu_baseline"]["kind"] == ("reference" if live.available() else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in line["config"]
    port = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1", "--cpu-port",
                           "--steps", "1", "--warmup", "0", "--cpu-rays", "40"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert json.loads(port.stdout.strip().splitlines()[-1])["cpu_baseline"]["kind"] == "port"


def test_device_array_numpy_semantics_on_cpu_tensor():
    """DeviceArray only needs a torch tensor: the lazy numpy view, slicing, in-place slices and arithmetic
    behave like the reference's numpy arrays (exercised here with a CPU tensor)."""
    import torch
    from turbulence_tracing_b200 import DeviceArray
    t = torch.arange(12, dtype=torch.float64).reshape(4, 3)
    a = DeviceArray(t)
    assert a.shape == (4, 3) and a.ndim == 2 and a.size == 12 and len(a) == 4 and a.dtype == np.float64
    np.testing.assert_array_equal(np.asarray(a), np.arange(12.0).reshape(4, 3))
    np.testing.assert_array_equal(a[1], [3.0, 4.0, 5.0])
    a[0:4:2, :] *= 1e3                                # the idiom of example_kitchensink.py:94
    assert a.torch[2, 1].item() == 7000.0 and np.asarray(a)[0, 2] == 2000.0
    assert isinstance(a * 2, np.ndarray) and (2 * a)[1, 1] == 8.0 and (a - a).sum() == 0 and (-a)[1, 0] == -3.0
    assert a.T.shape == (3, 4) and a.max() == 8000.0 and a.perm is None
    assert "DeviceArray" in repr(a)


def test_optics_program_composition():
    from turbulence_tracing_b200 import ray_transfer_matrix as rtm, _lib
    p0 = rtm.OpticsProgram()
    p1 = p0.distance(10).sym_lens(5)
    p2 = p1.angular_filter([0, 1, 2, 3]).knife_edge(0.5, "x", -1)
    assert len(p0) == 0 and len(p1) == 2 and len(p2) == 5          # immutable builder
    assert p2.ops[0] == (_lib.OP_DISTANCE, 10.0, 0.0) and p2.ops[1] == (_lib.OP_LENS, 5.0, 5.0)
    assert p2.ops[2] == (_lib.OP_ANNULAR_STOP, 0.0, 1.0) and p2.ops[3] == (_lib.OP_ANNULAR_STOP, 2.0, 3.0)
    assert p2.ops[4] == (_lib.OP_KNIFE_EDGE, 0.5, -1.0)
    assert rtm.OpticsProgram().knife_edge(0.0, "y", 1).ops[0][2] == 2.0
    with pytest.raises(ValueError):
        rtm.OpticsProgram().knife_edge(0.0, "z", 1)
