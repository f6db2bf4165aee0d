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
    # every parameter starts on a 16-byte boundary of the flat buffer
    assert b.flat.numel() == sum((p.numel() + 3) // 4 * 4 for p in m.parameters())
    assert all(o % 4 == 0 for o in b.offsets)
    assert m[1].bias.grad.data_ptr() == b.flat.data_ptr() + 4 * b.offsets[3]
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


def _dp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pygho_b200.dist import shard_contiguous
    torch.manual_seed(100 + rank)                                # replicas start different ...
    m = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    broadcast_parameters(m)                                      # ... and are made identical
    b = FlatGradBucket(m.parameters())
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    g = torch.Generator().manual_seed(7)
    X, y = torch.randn(16, 6, generator=g), torch.randn(16, 1, generator=g)
    mine = list(shard_contiguous(16, rank, world))               # whole "graphs" per rank
    for _ in range(3):
        b.zero()
        torch.nn.functional.mse_loss(m(X[mine]), y[mine]).backward()
        b.allreduce_mean()
        opt.step()
    out[rank] = torch.cat([p.detach().flatten() for p in m.parameters()]).tolist()
    dist.destroy_process_group()


def test_gloo_world2_data_parallel_equals_single_process():
    """Sharding the batch over two ranks + FlatGradBucket.allreduce_mean reproduces
    single-process training on the whole batch (equal shard sizes: mean of means)."""
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_dp_worker, args=(2, port, out), nprocs=2, join=True)
    torch.manual_seed(100)                                       # rank 0's initial weights
    m = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    g = torch.Generator().manual_seed(7)
    X, y = torch.randn(16, 6, generator=g), torch.randn(16, 1, generator=g)
    for _ in range(3):
        opt.zero_grad()
        torch.nn.functional.mse_loss(m(X), y).backward()
        opt.step()
    want = torch.cat([p.detach().flatten() for p in m.parameters()])
    assert torch.allclose(torch.tensor(out[0]), want, atol=1e-6)
    assert out[0] == out[1]                                      # replicas stay bit-identical
