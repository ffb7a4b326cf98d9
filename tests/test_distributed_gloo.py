"""world_size-2 tests of the multi-GPU host logic on CPU with the gloo backend: the flat gradient
bucket's single all-reduce (data-parallel replicas) and the block sharding of inference."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pointcloud_bridge_b200 import distributed as pdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, w, _ = pdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world) and pdist.is_dist()
    torch.manual_seed(0)                                   # identical replicas
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.ReLU(), torch.nn.Linear(3, 2))
    bucket = pdist.FlatGradBucket(net)
    assert bucket.flat.numel() == sum(p.numel() for p in net.parameters())
    # every rank trains on its own shard of one global batch of 8 rows
    torch.manual_seed(1)
    x, y = torch.randn(8, 4), torch.randn(8, 2)
    sl = pdist.shard_range(8, rank, world)
    bucket.zero()
    assert all(p.grad is None for p in net.parameters())
    loss = ((net(x[sl.start:sl.stop]) - y[sl.start:sl.stop]) ** 2).mean()
    loss.backward()
    bucket.pack()                                          # one batched copy into the flat buffer
    bucket.allreduce_mean()                                # one collective
    bucket.unpack()
    # reference: full-batch gradient on one process (equal shard sizes -> mean of shard means)
    torch.manual_seed(0)
    ref = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.ReLU(), torch.nn.Linear(3, 2))
    ((ref(x) - y) ** 2).mean().backward()
    ok = all(torch.allclose(p.grad, q_.grad, atol=1e-6) for p, q_ in zip(net.parameters(), ref.parameters()))
    t = pdist.max_over_ranks(float(rank + 1), "cpu")
    q.put((rank, ok, t, list(sl)))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_bucket_allreduce_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]
    assert [r[2] for r in res] == [2.0, 2.0]               # max over ranks
    assert res[0][3] == [0, 1, 2, 3] and res[1][3] == [4, 5, 6, 7]


@pytest.mark.parametrize("n,world", [(12208, 8), (10, 4), (3, 8), (0, 2)])
def test_shard_range_partitions_blocks(n, world):
    parts = [pdist.shard_range(n, r, world) for r in range(world)]
    flat = [i for p in parts for i in p]
    assert flat == list(range(n))                          # contiguous, disjoint, complete
    assert max(len(p) for p in parts) == -(-n // world) if n else True
