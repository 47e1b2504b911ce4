"""Rectilinear (stretched) mesh at scale: 257^3 nodes, 1e7 rays -- event marching (variant 0) vs the gather
kernel alone (variant 2), float32 and float64 grids.  Prints ray-steps/s; `--one` runs a single event-marching
solve (for ncu)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbulence_tracing_b200 import particle_tracker as pt, turboGen as tg
M, NR = 257, 10_000_000
t = np.linspace(-1, 1, M)
x = 5e-3 * np.tanh(1.6 * t) / np.tanh(1.6)                 # fine in the middle, coarse outside (ratio ~6)
y = 5e-3 * np.sign(t) * np.abs(t) ** 1.3
z = -5e-3 + 1e-2 * np.linspace(0, 1, M) ** 1.5             # fine at the entry face
f = tg.gaussian3D_FFT(128, lambda k: k ** (-11.0 / 3.0), seed=1234, dtype="float32", return_device=True).torch
ne = 1e25 * torch.clamp(1 + 0.3 * f / f.std(), min=0); del f
one = "--one" in sys.argv
for dtype, variant in ((("float32", 0),) if one else (("float32", 0), ("float64", 0), ("float32", 2), ("float64", 2))):
    cube = pt.ElectronCube(x, y, z, dtype=dtype, steps_per_cell=1, verbose=False, keep_sf=False)
    cube.external_ne(ne); cube.calc_dndr()
    assert cube._nodes is not None
    cube.kernel_variant = variant
    cube.init_beam(NR, 4e-3, 0.05e-3, seed=99)
    cube.solve()
    if one:
        break
    cube._trace_events = []
    for _ in range(3):
        cube.solve()
    torch.cuda.synchronize()
    ms = float(np.mean(cube.trace_ms()))
    ok = int((cube.status.torch == 1).sum())
    print(f"{dtype} grid, variant {variant}: trace {ms:.2f} ms, {cube.ray_steps / ms * 1e3:.3e} ray-steps/s, {ok} of {NR} rays through the far face")
