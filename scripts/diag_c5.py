"""1025^3 cube, 1.25e8 rays: where do the rays that miss the shadowgraphy image go?"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbulence_tracing_b200 import particle_tracker as pt, ray_transfer_matrix as rtm, turboGen as tg
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
M = 2 * N + 1
x = np.linspace(-5e-3, 5e-3, M)
f = tg.gaussian3D_FFT(N, lambda k: k ** (-11.0 / 3.0), seed=1234, dtype="float32", return_device=True).torch
ne = 1e25 * torch.clamp(1 + 0.3 * f / f.std(), min=0)
del f
cube = pt.ElectronCube(x, x, x, keep_sf=False, verbose=False)
cube.external_ne(ne); cube.calc_dndr()
cube.init_beam(125_000_000, 4e-3, 0.05e-3, seed=99)
rf = cube.solve().torch
st = cube.status.torch
print("status counts:", {int(v): int(c) for v, c in zip(*torch.unique(st, return_counts=True))})
th = torch.sqrt(rf[1] ** 2 + rf[3] ** 2)
print("rms angle %.3f mrad, max %.3f mrad, rays > 50 mrad: %d, > 62.5 mrad: %d, non-finite: %d" % (
    float(th.pow(2).mean().sqrt()) * 1e3, float(th.max()) * 1e3, int((th > 0.05).sum()), int((th > 0.0625).sum()),
    int((~torch.isfinite(rf)).any(0).sum())))
sh = rtm.Shadowgraphy(cube.rf); sh.solve(); d = sh.rf.torch
print("rejected by the lens apertures:", int(torch.isnan(d[0]).sum()), " on detector:", end=" ")
sh.histogram(); print(int(sh.H.sum()))
