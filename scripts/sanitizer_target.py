#!/usr/bin/env python
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): every kernel of the hot path once, with
the edge cases that take the rare branches -- a wide divergent beam (side exits, misses: deferred rays, compacted list,
general kernel), Morton sort, face-coefficient path AND corner-grid path, FP64, optics + privatised histogram at two
binnings, the passive quantities, a chunked host upload with a 1-ray tail."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from turbulence_tracing_b200 import particle_tracker as pt, ray_transfer_matrix as rtm, turboGen as tg
import torch

f = tg.gaussian3D_FFT(20, lambda k: k ** (-11.0 / 3.0), seed=5, dtype="float32", return_device=True).torch
ne = 1e25 * torch.clamp(1 + 0.3 * f / f.std(), min=0)
M = ne.shape[0]
x = np.linspace(-5e-3, 5e-3, M)
tot = 0
for dtype, fg, spc in (("float32", "auto", 1), ("float32", False, 1), ("float32", False, 3), ("float64", False, 2)):
    cube = pt.ElectronCube(x, x, x, dtype=dtype, steps_per_cell=spc, verbose=False, face_grid=fg)
    cube.external_ne(ne)
    cube.calc_dndr()
    cube.init_beam(20_000, 5.5e-3, 2e-2, seed=3)           # wide and divergent
    rf = cube.solve()
    st = cube.status.torch
    assert 0 < int((st != 1).sum()) < 20_000
    for bs in (10, 1):
        sh = rtm.Shadowgraphy(rf); sh.solve(); sh.histogram(bin_scale=bs)
        tot += sh.H.sum()
    sc = rtm.Schlieren_DF(rf); sc.solve(R=1); sc.histogram()
    s0 = np.asarray(cube.s0)[:, :4001].copy()              # host rays, chunked upload, 1-ray tail
    cube.s0 = s0
    cube.pipeline_first_rays = 500
    cube.pipeline_chunk_rays = 1000
    cube.solve()
cube = pt.ElectronCube(x, x, x, B_on=True, inv_brems=True, phaseshift=True, verbose=False)
cube.external_ne(ne)
B = torch.zeros((M, M, M, 3), dtype=torch.float32, device="cuda"); B[..., 2] = 5.0
cube.external_B(B); cube.external_Te(torch.full((M, M, M), 100.0, device="cuda")); cube.external_Z(1.0)
cube.calc_dndr()
cube.init_beam(5000, 4e-3, 1e-3, seed=4)
cube.solve()
# round 2, session 3: the privatised detector kernel (>= 65536 rays in storage order; even and odd counts = double2 / scalar
# loads; a focused bundle that overflows the 16-bit counters), the fused aux grid with cubes for Te and Z in FP32 and
# FP64, the staged upload of a pageable array (more pieces than ring slots)
rng = np.random.default_rng(1)
for n in (70_000, 70_001):
    r0 = np.zeros((4, n)); r0[0] = rng.uniform(-10e-3, 10e-3, n); r0[2] = rng.uniform(-7e-3, 7e-3, n); r0[1, ::7] = np.nan
    sh = rtm.Shadowgraphy(r0); sh.solve(); sh.histogram()
    tot += sh.H.sum()
r0 = np.zeros((4, 5_200_000)); r0[0] = 1e-4; r0[2] = 2e-4
sh = rtm.Shadowgraphy(r0); sh.solve(); sh.histogram()
assert sh.H.max() == 5_200_000
for dt in (torch.float32, torch.float64):
    c2 = pt.ElectronCube(x, x, x, "y", B_on=True, inv_brems=True, phaseshift=True, verbose=False, dtype=str(dt).replace("torch.", ""))
    c2.external_ne(ne.to(dt)); c2.external_B(B.to(dt)); c2.external_Te(torch.full((M, M, M), 80.0, device="cuda", dtype=dt))
    c2.external_Z(torch.full((M, M, M), 2.0, device="cuda", dtype=dt)); c2.calc_dndr(); c2.set_up_interps()
from turbulence_tracing_b200 import _lib
a = torch.from_numpy(rng.integers(0, 1 << 40, size=(150 << 20) // 8 + 3))
d = _lib.h2d(None, a)
torch.cuda.synchronize()
assert torch.equal(d.cpu(), a)
torch.cuda.synchronize()
print("sanitizer target done", M, int(tot))
