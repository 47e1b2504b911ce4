"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: sharding of one global beam, broadcast of
the cube, and the single all-reduce of concatenated integer histograms (N-rank image == 1-rank image)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from turbulence_tracing_b200 import distributed as ttd


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 100, 10**8 + 3):
        for world in (1, 2, 3, 8):
            parts = [ttd.shard_range(n, r, world) for r in range(world)]
            assert sum(c for _, c in parts) == n
            assert parts[0][0] == 0
            for (f0, c0), (f1, _) in zip(parts, parts[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    with pytest.raises(ValueError):
        ttd.shard_range(10, 2, 2)


def test_upload_cube_sharded_single_process():
    a = np.arange(24, dtype=np.float64).reshape(2, 3, 4)
    t = ttd.upload_cube_sharded(a, device=torch.device("cpu"))
    assert t.dtype == torch.float64 and np.array_equal(t.numpy(), a)


def test_single_process_collectives_are_identity():
    h = [torch.arange(6).reshape(2, 3), torch.ones(4, dtype=torch.int64)]
    out = ttd.allreduce_histograms(h)
    assert all(torch.equal(a.to(torch.int64), b) for a, b in zip(h, out))
    assert ttd.allreduce_scalar(5) == 5 and ttd.rank_world() == (0, 1)
    t = torch.zeros(3)
    assert ttd.broadcast_cube(t) is t


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from oracle import ref_numpy as orc
    r, w, _ = ttd.init_from_env(backend="gloo")
    assert (r, w) == (rank, world) and ttd.rank_world() == (rank, world)
    # cube broadcast from rank 0
    ne = torch.full((5, 5, 5), float(rank + 1))
    if rank == 0:
        ne = torch.arange(125, dtype=torch.float32).reshape(5, 5, 5)
    ttd.broadcast_cube(ne, src=0)
    assert torch.equal(ne, torch.arange(125, dtype=torch.float32).reshape(5, 5, 5))
    # host cube -> every rank's device: each rank uploads its slice, one all-gather (odd size: the last slice is short)
    rng0 = np.random.RandomState(11)
    host = rng0.rand(7, 5, 3).astype(np.float32)
    mine = host.copy()
    lo = (rank * 53) % host.size                    # whatever lies outside this rank's slice is never read
    got = ttd.upload_cube_sharded(mine, device=torch.device("cpu"))
    assert got.shape == host.shape and got.dtype == torch.float32 and np.array_equal(got.numpy(), host)
    # one global bundle of detector-plane rays, sharded; per-rank histograms from the oracle
    rng = np.random.RandomState(3)
    n = 5001
    rf = np.zeros((4, n))
    rf[0], rf[2] = rng.uniform(-10, 10, n), rng.uniform(-8, 8, n)
    first, count = ttd.shard_range(n, rank, world)
    H1, _, _ = orc.histogram(rf[:, first:first + count])
    H2, _, _ = orc.histogram(rf[:, first:first + count], Lx=6, Ly=6, bin_scale=25)
    red = ttd.allreduce_histograms([torch.from_numpy(H1).to(torch.int64), torch.from_numpy(H2).to(torch.int64)])
    F1, _, _ = orc.histogram(rf)
    F2, _, _ = orc.histogram(rf, Lx=6, Ly=6, bin_scale=25)
    assert np.array_equal(red[0].numpy(), F1) and np.array_equal(red[1].numpy(), F2)
    assert ttd.allreduce_scalar(count, "sum") == n
    assert ttd.allreduce_scalar(float(rank), "max") == world - 1
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
    open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")


@pytest.mark.timeout(300)
def test_two_rank_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
