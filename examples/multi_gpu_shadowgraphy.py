"""The reference's example_MPI.py pattern (particle_tracking/example_MPI.py:82-166) on the B200 backend:
every rank holds the cube, traces its own shard of one global beam in bundles, the detector images are
all-reduced once.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        examples/multi_gpu_shadowgraphy.py 1e9
(also runs on a single GPU without torchrun)
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbulence_tracing_b200 import distributed as ttd                                        # noqa: E402
from turbulence_tracing_b200 import particle_tracker as pt, ray_transfer_matrix as rtm, turboGen as tg  # noqa: E402

Np = int(float(sys.argv[1])) if len(sys.argv) > 1 else int(1e7)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 128
rank, world, local = ttd.init_from_env()
M = 2 * N + 1
x = np.linspace(-5e-3, 5e-3, M)

ne = torch.empty((M, M, M), dtype=torch.float32, device="cuda")
if rank == 0:                                   # example_MPI.py:97-111 rebuilt the cube on every rank
    f = tg.gaussian3D_FFT(N, lambda k: k ** (-11.0 / 3.0), seed=1234, dtype="float32", return_device=True).torch
    ne.copy_(1e25 * torch.clamp(1 + 0.3 * f / f.std(), min=0))
ttd.broadcast_cube(ne, src=0)

cube = pt.ElectronCube(x, x, x, "z", keep_sf=False, verbose=False)
cube.external_ne(ne)
cube.calc_dndr()
detectors = [(rtm.Shadowgraphy, {}, {}), (rtm.Schlieren_DF, {}, {"R": 1}), (rtm.Refractometer, {}, {})]
H, ray_steps = ttd.trace_sharded(cube, detectors, Np, beam_size=4e-3, divergence=0.05e-3, seed=1,
                                 bundle=int(5e7))                  # Np_ray_split, example_MPI.py:86
if rank == 0:
    print(f"world {world}: {Np} rays, {ray_steps} ray-steps; counts per detector "
          f"{[int(h.sum()) for h in H]}; image shape {tuple(H[0].shape)}")
if world > 1:
    torch.distributed.destroy_process_group()
