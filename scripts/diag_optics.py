"""optics+hist kernel: Morton-ordered gather (perm) vs original-order streaming, 1e8 rays"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbulence_tracing_b200 import particle_tracker as pt, ray_transfer_matrix as rtm, turboGen as tg
M = 257
x = np.linspace(-5e-3, 5e-3, M)
f = tg.gaussian3D_FFT(128, lambda k: k ** (-11.0 / 3.0), seed=1234, dtype="float32", return_device=True).torch
ne = 1e25 * torch.clamp(1 + 0.3 * f / f.std(), min=0)
cube = pt.ElectronCube(x, x, x, keep_sf=False, verbose=False)
cube.external_ne(ne); cube.calc_dndr()
cube.init_beam(100_000_000, 4e-3, 0.05e-3, seed=99)
rf = cube.solve()
for name, perm in (("perm (Morton order)", rf.perm), ("original order", None)):
    for bs in (10, 1):
        ts = []
        for it in range(4):
            sh = rtm.Shadowgraphy(rf); sh._perm = perm; sh.solve()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); sh.histogram(bin_scale=bs, to_host=False); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(f"{name:22s} bin_scale={bs:2d}: {min(ts):7.2f} ms   sum={int(sh.H_dev.sum())}")
