"""2-rank diagnostic: is the 34 ms 'allreduce' phase NCCL cost or rank skew?"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbulence_tracing_b200 import distributed as ttd, particle_tracker as pt, turboGen as tg
rank, world, local = ttd.init_from_env()
dev = torch.device("cuda", local)
h = torch.ones((257, 344), dtype=torch.int64, device=dev)
for _ in range(3):
    ttd.allreduce_histograms([h])
torch.cuda.synchronize(); torch.distributed.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ttd.allreduce_histograms([h])
e1.record(); torch.cuda.synchronize()
print(f"rank {rank}: bare allreduce {e0.elapsed_time(e1) / 20:.3f} ms", flush=True)
# per-rank trace time of identical work
M = 257
x = np.linspace(-5e-3, 5e-3, M)
f = tg.gaussian3D_FFT(128, lambda k: k ** (-11.0 / 3.0), seed=1234, dtype="float32", return_device=True).torch
ne = 1e25 * torch.clamp(1 + 0.3 * f / f.std(), min=0)
cube = pt.ElectronCube(x, x, x, keep_sf=False, verbose=False)
cube.external_ne(ne); cube.calc_dndr()
cube.init_beam(20_000_000, 4e-3, 0.05e-3, seed=99, first_ray=0)   # SAME rays on both ranks
for it in range(4):
    torch.cuda.synchronize(); torch.distributed.barrier()
    cube._trace_events = []
    t0 = time.perf_counter()
    cube.solve()
    torch.cuda.synchronize()
    print(f"rank {rank} it {it}: trace kernel {cube.trace_ms()[0]:.2f} ms, solve wall {1e3 * (time.perf_counter() - t0):.2f} ms", flush=True)
torch.distributed.barrier(); torch.distributed.destroy_process_group()
