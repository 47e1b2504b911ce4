"""Where does the e2e step spend its time?  513^3, 1e8 rays."""
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
ne_pin = torch.empty((M, M, M), dtype=torch.float32, pin_memory=True); ne_pin.copy_(ne)
s0_np, ne_np = s0_pin.numpy(), ne_pin.numpy()
def t(fn, n=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return 1e3 * (time.perf_counter() - t0) / n
def solve_dev(): cube.s0 = s0_dev; cube.solve()
def solve_pin(): cube.s0 = s0_np; cube.solve()
def h2d_only(): torch.from_numpy(s0_np).cuda()
def dndr_host(): cube.external_ne(ne_np); cube.calc_dndr()
def dndr_dev(): cube.external_ne(ne); cube.calc_dndr()
print(f"solve, device rays        {t(solve_dev):8.1f} ms")
print(f"solve, pinned host rays   {t(solve_pin):8.1f} ms   (is_pinned={torch.from_numpy(s0_np).is_pinned()})")
print(f"plain H2D of the rays     {t(h2d_only):8.1f} ms   ({s0_np.nbytes / 1e9:.1f} GB)")
print(f"calc_dndr, host cube      {t(dndr_host):8.1f} ms   calc_dndr, device cube {t(dndr_dev):8.1f} ms")
for chunk in (6_250_000, 25_000_000):
    cube.pipeline_chunk_rays = chunk
    print(f"solve, pinned, chunk {chunk:>9d}: {t(solve_pin):8.1f} ms")

# the bench's e2e step, piece by piece
cube.pipeline_chunk_rays = 12_500_000
cube2 = pt.ElectronCube(x, x, x, keep_sf=False, verbose=False)
def e2e_step():
    cube2.external_ne(ne_np); cube2.calc_dndr()
    cube2.s0 = s0_np
    rf = cube2.solve()
    sh = rtm.Shadowgraphy(rf); sh.solve(); sh.histogram()
    _ = cube2.ray_steps
print(f"bench-like e2e step       {t(e2e_step):8.1f} ms")
def e2e_nosync():
    cube2.external_ne(ne_np); cube2.calc_dndr()
    cube2.s0 = s0_np
    rf = cube2.solve()
    sh = rtm.Shadowgraphy(rf); sh.solve(); sh.histogram(to_host=False)
print(f"same without D2H syncs    {t(e2e_nosync):8.1f} ms")
import gc
def e2e_fresh():
    c3 = pt.ElectronCube(x, x, x, keep_sf=False, verbose=False)
    c3.external_ne(ne_np); c3.calc_dndr(); c3.s0 = s0_np
    rf = c3.solve(); sh = rtm.Shadowgraphy(rf); sh.solve(); sh.histogram(); _ = c3.ray_steps
print(f"fresh cube object per step{t(e2e_fresh):8.1f} ms")
print(torch.cuda.memory_summary(abbreviated=True)[:1500])
