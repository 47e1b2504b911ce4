"""What happens between the end of solve() and the end of the e2e step (bench.py: 'optics+hist+d2h')?"""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbulence_tracing_b200 import particle_tracker as pt, ray_transfer_matrix as rtm, turboGen as tg
M = 513
NR = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
x = np.linspace(-5e-3, 5e-3, M)
f = tg.gaussian3D_FFT(256, lambda k: k ** (-11.0 / 3.0), seed=1234, dtype="float32", return_device=True).torch
ne = 1e25 * torch.clamp(1 + 0.3 * f / f.std(), min=0); del f
cube = pt.ElectronCube(x, x, x, keep_sf=False, verbose=False)
cube.external_ne(ne); cube.calc_dndr()
cube.init_beam(NR, 4e-3, 0.05e-3, seed=99)
s0_dev = cube.s0
s0_pin = torch.empty((6, NR), dtype=torch.float64, pin_memory=True); s0_pin.copy_(s0_dev.torch)
s0_np = s0_pin.numpy()

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

for mode in ("device", "host", "host"):
    cube.s0 = s0_dev if mode == "device" else s0_np
    torch.cuda.synchronize()
    h0 = time.perf_counter(); e0 = ev()
    rf = cube.solve()
    h1 = time.perf_counter(); e1 = ev()
    sh = rtm.Shadowgraphy(rf); sh.solve()
    h2 = time.perf_counter(); e2 = ev()
    sh.histogram(to_host=False)
    h3 = time.perf_counter(); e3 = ev()
    H = sh.H_dev.double().cpu().numpy()
    h4 = time.perf_counter(); e4 = ev()
    n = cube.ray_steps
    h5 = time.perf_counter(); e5 = ev()
    torch.cuda.synchronize()
    g = [a.elapsed_time(b) for a, b in ((e0, e1), (e1, e2), (e2, e3), (e3, e4), (e4, e5))]
    h = [1e3 * (b - a) for a, b in ((h0, h1), (h1, h2), (h2, h3), (h3, h4), (h4, h5))]
    print(f"{mode:6s} gpu ms: solve {g[0]:7.1f} | Shadowgraphy+solve {g[1]:6.2f} | histogram {g[2]:6.2f} | H d2h {g[3]:6.2f} | ray_steps {g[4]:6.2f}")
    print(f"{'':6s} host ms: solve {h[0]:7.1f} | Shadowgraphy+solve {h[1]:6.2f} | histogram {h[2]:6.2f} | H d2h {h[3]:6.2f} | ray_steps {h[4]:6.2f}")
