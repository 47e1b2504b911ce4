"""e2e solve() with pinned host rays: chunk cap x growth factor of the upload pipeline.  513^3, 1e8 rays."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbulence_tracing_b200 import particle_tracker as pt, ray_transfer_matrix as rtm, turboGen as tg
M = 513
x = np.linspace(-5e-3, 5e-3, M)
f = tg.gaussian3D_FFT(256, lambda k: k ** (-11.0 / 3.0), seed=1234, dtype="float32", return_device=True).torch
ne = 1e25 * torch.clamp(1 + 0.3 * f / f.std(), min=0); del f
cube = pt.ElectronCube(x, x, x, keep_sf=False, verbose=False)
cube.external_ne(ne); cube.calc_dndr()
cube.init_beam(100_000_000, 4e-3, 0.05e-3, seed=99)
s0_dev = cube.s0
s0_pin = torch.empty((6, 100_000_000), dtype=torch.float64, pin_memory=True); s0_pin.copy_(s0_dev.torch)
s0_np = s0_pin.numpy()
def times(fn, n=4):
    out = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        out.append(round(1e3 * (time.perf_counter() - t0), 1))
    return out
def solve_dev(): cube.s0 = s0_dev; cube.solve()
def solve_pin(): cube.s0 = s0_np; cube.solve()
def full():
    cube.s0 = s0_np; rf = cube.solve(); sh = rtm.Shadowgraphy(rf); sh.solve(); sh.histogram(); _ = cube.ray_steps
print("device rays", times(solve_dev))
for cap, g in ((12_500_000, 2), (25_000_000, 2), (50_000_000, 2), (50_000_000, 3), (100_000_000, 3), (100_000_000, 4)):
    cube.pipeline_chunk_rays, cube.pipeline_growth = cap, g
    cube._trace_events = []
    a = times(solve_pin)
    ev = cube._trace_events; cube._trace_events = None
    k = len(ev) // 4
    tr = sum(e0.elapsed_time(e1) for e0, e1 in ev[-k:])
    print(f"cap {cap:>9d} growth {g}: solve {a} ms; {k} trace launches, kernels {tr:.1f} ms; solve+optics+hist+d2h {times(full)}")
