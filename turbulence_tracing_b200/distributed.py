"""Multi-GPU driver: one process per GPU (torchrun), NCCL over NVLink.

This is the B200 equivalent of the reference's MPI pattern (particle_tracking/example_MPI.py:82-149):
every rank holds the whole cube, traces its own bundle of rays, and the detector images are
sum-reduced.  Here

  * the ne cube is broadcast from rank 0 (``broadcast_cube``; the reference rebuilds it on every
    rank, example_MPI.py:97-111), each rank runs ``calc_dndr`` locally;
  * rays are never exchanged: rank r generates ray ids ``shard_range(Np, r, world)`` of ONE global
    beam with the counter-based device RNG (``ElectronCube.init_beam(seed=, first_ray=)``), so the
    union over ranks is the same beam for any world size;
  * the histograms of all detectors are concatenated into one int64 buffer and all-reduced once
    (``allreduce_histograms``; the reference does three pickled ``comm.reduce`` calls,
    example_MPI.py:147-149).  Integer counts make the N-GPU image equal the 1-GPU image exactly.

The collective payload is tiny (354 KB per detector at the default binning), so there is no fused
compute+collective kernel here: nothing to overlap.  Works with the ``gloo`` backend on CPU tensors
too (used by the CPU tests of this logic).
"""
from __future__ import annotations

import os


def _dist():
    import torch.distributed as dist
    return dist


def is_initialized() -> bool:
    dist = _dist()
    return dist.is_available() and dist.is_initialized()


def init_from_env(backend: str | None = None):
    """Initialise torch.distributed from torchrun's environment (RANK, WORLD_SIZE, LOCAL_RANK,
    MASTER_ADDR, MASTER_PORT) and bind this process to its GPU.  Returns (rank, world, local_rank)."""
    import torch
    dist = _dist()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def bind_to_gpu_numa_node(local_rank: int):
    """Pin this process (and the host buffers it first-touches from now on) to the CPUs of the NUMA node its GPU hangs
    off: /sys/bus/pci/devices/<gpu>/local_cpulist.  With 8 ranks uploading 5 GB of launch rays each per step, a staging
    buffer on the far socket halves the host->device rate.  Best effort: returns the CPU list used, or None."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return spec
    except Exception:
        return None


def rank_world():
    if is_initialized():
        dist = _dist()
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous, balanced split of ray ids [0, n_total): returns (first, count).  The first
    ``n_total % world`` ranks get one extra ray; an empty shard is legal."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("need 0 <= rank < world")
    base, extra = divmod(int(n_total), world)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def broadcast_cube(ne, src: int = 0):
    """Broadcast the ne cube (torch tensor, same shape/dtype allocated on every rank) from ``src``.
    In place; returns the tensor.  No-op for a single process."""
    if is_initialized() and _dist().get_world_size() > 1:
        _dist().broadcast(ne, src=src)
    return ne


def upload_cube_sharded(ne_host, device=None):
    """Host cube -> device cube on EVERY rank, each rank uploading only its 1/world slice over PCIe and the slices
    exchanged with one all-gather over NVLink.  The reference's MPI pattern has every rank build / load the whole cube
    (example_MPI.py:97-111); with 8 ranks sharing one host that is 8 x 540 MB of host reads per 513^3 cube (16-27 ms per
    rank, measured) against 67 MB + a sub-millisecond collective.  ``ne_host``: C-contiguous numpy array, the same on
    every rank (only this rank's slice is read).  Returns a torch tensor of the same shape and dtype on ``device``.
    Single process: a plain upload."""
    import numpy as np
    import torch
    a = np.ascontiguousarray(ne_host)
    flat = a.reshape(-1)
    dt = torch.from_numpy(flat[:0]).dtype
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    rank, world = rank_world()
    from . import _lib
    on_gpu = torch.device(device).type == "cuda"

    def up(dst, src):                    # pageable sources go through the staging threads of tt_h2d_pageable
        if on_gpu:
            with torch.cuda.device(dst.device):
                _lib.h2d(dst, src)
        else:
            dst.copy_(src)

    if world == 1:
        out = torch.empty(a.shape, dtype=dt, device=device)
        up(out.view(-1), torch.from_numpy(flat))
        return out
    n = flat.size
    chunk = (n + world - 1) // world
    out = torch.empty(world * chunk, dtype=dt, device=device)
    lo, hi = min(rank * chunk, n), min((rank + 1) * chunk, n)
    mine = out[rank * chunk:(rank + 1) * chunk]
    if hi > lo:
        up(mine[:hi - lo], torch.from_numpy(flat[lo:hi]))
    if hi - lo < chunk:
        mine[hi - lo:].zero_()
    _dist().all_gather_into_tensor(out, mine)          # in place: rank r's input is slice r of the output
    return out[:n].view(a.shape)


def allreduce_histograms(hists):
    """Sum-reduce a list of integer histograms (torch tensors on the backend's device) across ranks
    with ONE collective; returns the list of reduced tensors (new storage)."""
    import torch
    if not hists:
        return []
    flat = torch.cat([h.reshape(-1).to(torch.int64) for h in hists])
    if is_initialized() and _dist().get_world_size() > 1:
        _dist().all_reduce(flat, op=_dist().ReduceOp.SUM)
    out, o = [], 0
    for h in hists:
        n = h.numel()
        out.append(flat[o:o + n].reshape(h.shape))
        o += n
    return out


def allreduce_scalar(value, op: str = "sum", device=None):
    """Reduce one python number over ranks (sum or max)."""
    import torch
    if not (is_initialized() and _dist().get_world_size() > 1):
        return value
    dist = _dist()
    dt = torch.float64 if isinstance(value, float) else torch.int64
    t = torch.tensor([value], dtype=dt, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM if op == "sum" else dist.ReduceOp.MAX)
    return t.item()


def trace_sharded(cube, detectors, n_total, beam_size, divergence, seed, bundle=None, histogram_kw=None,
                  rank=None, world=None, reduce=True):
    """The reference's ``full_system_solve`` loop (example_MPI.py:48-79, 117-141) on this rank's shard
    of one global beam, followed by the single histogram all-reduce.

    cube: ElectronCube with its gradient grid built.  detectors: list of (cls, ctor_kwargs,
    solve_kwargs).  bundle: rays per launch (memory knob, the reference's Np_ray_split).
    rank/world override the process group's (to compute one shard's contribution in a single process;
    combine with reduce=False).
    Returns (list of reduced int64 histograms [device], total ray-steps over all ranks)."""
    import torch
    if rank is None or world is None:
        rank, world = rank_world()
    first, count = shard_range(n_total, rank, world)
    bundle = int(bundle or max(count, 1))
    histogram_kw = histogram_kw or {}
    acc = None
    steps = 0
    for lo in range(0, max(count, 0), bundle):
        n = min(bundle, count - lo)
        cube.init_beam(n, beam_size, divergence, seed=seed, first_ray=first + lo)
        rf = cube.solve()
        steps += cube.ray_steps
        hs = []
        for cls, ckw, skw in detectors:
            d = cls(rf, **ckw)
            d.solve(**skw)
            d.histogram(**histogram_kw)
            hs.append(d.H_dev)
        acc = hs if acc is None else [a + h for a, h in zip(acc, hs)]
    if acc is None:          # empty shard: contribute zeros of the right shape
        acc = []
        for cls, ckw, skw in detectors:
            pix_x, pix_y = histogram_kw.get("pix_x", 3448), histogram_kw.get("pix_y", 2574)
            bs = histogram_kw.get("bin_scale", 10)
            acc.append(torch.zeros((pix_y // bs, pix_x // bs), dtype=torch.int64, device="cuda"))
    if not reduce:
        return acc, int(steps)
    reduced = allreduce_histograms(acc)
    total_steps = allreduce_scalar(int(steps), "sum", device=reduced[0].device if reduced else None)
    return reduced, total_steps
