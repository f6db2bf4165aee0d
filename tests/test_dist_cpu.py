"""Host-side data-parallel logic on CPU: sharding and the flat gradient bucket with a
world_size-2 gloo group."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pygho_b200.dist import FlatGradBucket, broadcast_parameters, shard_by_cost, shard_contiguous


def test_shard_by_cost_balances_and_covers():
    costs = [5, 9, 1, 7, 3, 3, 8, 2]
    parts = shard_by_cost(costs, 3)
    assert sorted(i for p in parts for i in p) == list(range(len(costs)))
    loads = [sum(costs[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= max(costs)
    assert shard_by_cost(costs, 3) == parts                     # deterministic


def test_shard_contiguous():
    got = [list(shard_contiguous(10, r, 4)) for r in range(4)]
    assert sum(got, []) == list(range(10)) and all(len(g) in (2, 3) for g in got)


def test_flat_bucket_views():
    m = torch.nn.Sequential(torch.nn.Linear(3, 4), torch.nn.Linear(4, 2))
    b = FlatGradBucket(m.parameters())
    m(torch.ones(5, 3)).sum().backward()
    assert b.flat.numel() == sum(p.numel() for p in m.parameters())
    assert float(b.flat.abs().sum()) > 0
    assert m[0].weight.grad.data_ptr() == b.flat.data_ptr()
    b.zero()
    assert float(m[1].bias.grad.abs().sum()) == 0.0


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(rank)                                      # different init per rank
    m = torch.nn.Linear(4, 3)
    broadcast_parameters(m)
    b = FlatGradBucket(m.parameters())
    x = torch.full((2, 4), float(rank + 1))
    m(x).sum().backward()
    local = b.flat.clone()
    b.allreduce_mean()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    ws = [torch.zeros_like(m.weight) for _ in range(world)]
    dist.all_gather(ws, m.weight.data)
    ok = torch.allclose(b.flat, sum(gathered) / world) and torch.equal(ws[0], ws[1])
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_gloo_world2_allreduce_mean():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]
