"""GPU tests of the round-2 training plumbing: ticketed BatchNorm reductions with device-side
row counts and in-place gradient accumulation, cross-rank (SyncBN) statistics, and CUDA-graph
replay of host-fed steps on capacity-padded batches (pygho_b200/static.py)."""
import copy
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def close(a, b, rtol):
    a, b = a.detach().double().cpu().numpy(), b.detach().double().cpu().numpy()
    assert a.shape == b.shape, (a.shape, b.shape)
    scale = max(1.0, float(np.abs(b).max()) if b.size else 1.0)
    err = float(np.abs(a - b).max(initial=0.0))
    assert err <= rtol * scale, (err, scale)


def _mlp(cin, cout, seed):
    from pygho_b200.honn.utils import MLP
    torch.manual_seed(seed)
    mlp = MLP(cin, cout, 2, True, norm="bn", act="silu", normparam=0.3).to(DEV)
    with torch.no_grad():
        for m in mlp.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.5, 0.5)
    return mlp


@pytest.mark.parametrize("rows,cin,cout", [(5000, 128, 128), (70001, 384, 128), (333, 16, 12)])
def test_direct_gradient_accumulation_equals_autograd(rows, cin, cout):
    """With ``p.grad`` preset (flat bucket) the fused backward adds its parameter gradients in
    place (cuBLAS beta = 1, accumulate flags of the kernels); the result equals autograd's own
    accumulation, for a second backward on top of a first one too."""
    from pygho_b200.dist import FlatGradBucket
    mlp = _mlp(cin, cout, rows)
    ref = copy.deepcopy(mlp)
    bucket = FlatGradBucket(mlp.parameters())        # p.grad = views of one buffer -> direct path
    x = torch.randn(rows, cin, device=DEV)
    w = torch.randn(rows, cout, device=DEV)
    for rep in range(2):                             # second pass accumulates on top
        (mlp(x) * w).sum().backward()
        from pygho_b200 import ops
        ops.set_direct_grad_accumulation(False)
        try:
            (ref(x) * w).sum().backward()
        finally:
            ops.set_direct_grad_accumulation(True)
        for (k, p), (_, q) in zip(mlp.named_parameters(), ref.named_parameters()):
            assert p.grad.data_ptr() >= bucket.flat.data_ptr()
            if k.endswith("bias") and "norm" not in k:      # exact zero + rounding noise
                assert float(p.grad.abs().max()) < 5e-3
                continue
            close(p.grad, q.grad, 2e-5)


@pytest.mark.parametrize("rows,cap,C", [(1000, 1100, 128), (4097, 4104, 384), (50, 900, 16)])
def test_bn_rows_dev_ignores_pad_rows(rows, cap, C):
    """Capacity-padded tensors + device row count == the exact-size tensors: statistics,
    running stats, outputs (pads -> 0), gradients (pads -> 0), parameter gradients."""
    ops = torch.ops.pygho_b200
    g = torch.Generator(device=DEV).manual_seed(rows)
    y = torch.randn(rows, C, device=DEV, generator=g) * 2 + 1
    dz = torch.randn(rows, C, device=DEV, generator=g)
    res = torch.randn(rows, C, device=DEV, generator=g)
    gamma = torch.rand(C, device=DEV, generator=g) + 0.5
    beta = torch.randn(C, device=DEV, generator=g)
    junk = lambda: torch.randn(cap - rows, C, device=DEV, generator=g) * 100  # noqa: E731
    yp, dzp, resp = torch.cat([y, junk()]), torch.cat([dz, junk()]), torch.cat([res, junk()])
    n = torch.tensor([rows], dtype=torch.int32, device=DEV)
    rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    rmp, rvp = rm.clone(), rv.clone()
    mean, rstd = ops.bn_stats(y, 1e-5, 0.1, rm, rv, None)
    meanp, rstdp = ops.bn_stats(yp, 1e-5, 0.1, rmp, rvp, n)
    assert torch.equal(rm, rmp) or float((rm - rmp).abs().max()) < 1e-6
    close(meanp, mean, 1e-6), close(rstdp, rstd, 1e-6), close(rvp, rv, 1e-6)
    z = ops.bn_act_fwd(y, mean, rstd, gamma, beta, 1, res, None)
    zp = ops.bn_act_fwd(yp, meanp, rstdp, gamma, beta, 1, resp, n)
    close(zp[:rows], z, 1e-6)
    assert float(zp[rows:].abs().max()) == 0.0
    sums, dg, db = ops.bn_act_bwd_reduce(dz, y, mean, rstd, gamma, beta, 1, None, None, None)
    sumsp, dgp, dbp = ops.bn_act_bwd_reduce(dzp, yp, meanp, rstdp, gamma, beta, 1, n, None, None)
    close(sumsp, sums, 1e-5), close(dgp, dg, 1e-5), close(dbp, db, 1e-5)
    dy, dbias = ops.bn_act_bwd_apply(dz, y, mean, rstd, gamma, beta, sums, None, 1, None, True, None)
    dyp, dbiasp = ops.bn_act_bwd_apply(dzp, yp, meanp, rstdp, gamma, beta, sumsp, None, 1, n, True, None)
    close(dyp[:rows], dy, 1e-5)
    assert float(dyp[rows:].abs().max()) == 0.0
    assert float(dbias.abs().max()) < 1e-2 and float(dbiasp.abs().max()) < 1e-2   # == 0 + noise
    # accumulate flags add to existing buffers
    acc_g, acc_b, acc_bias = torch.ones(C, device=DEV), torch.ones(C, device=DEV), torch.ones(C, device=DEV)
    s2, e1, e2 = ops.bn_act_bwd_reduce(dz, y, mean, rstd, gamma, beta, 1, None, acc_g, acc_b)
    assert e1.numel() == 0 and e2.numel() == 0 and torch.equal(s2, sums)
    close(acc_g, dg + 1, 1e-6), close(acc_b, db + 1, 1e-6)
    ops.bn_act_bwd_apply(dz, y, mean, rstd, gamma, beta, sums, None, 1, None, True, acc_bias)
    close(acc_bias, dbias + 1, 1e-5)
    # deterministic: the ticketed combine adds in a fixed order whichever CTA finishes last
    for _ in range(3):
        m2, r2 = ops.bn_stats(y, 1e-5, 0.1, None, None, None)
        assert torch.equal(m2, mean) and torch.equal(r2, rstd)
        s3, _a, _b = ops.bn_act_bwd_reduce(dz, y, mean, rstd, gamma, beta, 1, None, None, None)
        assert torch.equal(s3, sums)


def test_bn_sync_statistics_equal_whole_batch():
    """Per-rank (mean, M2, count) triples merged by bn_sync_finalize == statistics of the
    concatenated rows (what a single process would compute): the SyncBN forward on one GPU."""
    ops = torch.ops.pygho_b200
    g = torch.Generator(device=DEV).manual_seed(3)
    C = 128
    parts = [torch.randn(r, C, device=DEV, generator=g) * s + o
             for r, s, o in ((3000, 1.0, 0.0), (517, 3.0, 5.0), (12001, 0.5, -2.0))]
    whole = torch.cat(parts)
    rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    mean, rstd = ops.bn_stats(whole, 1e-5, 0.1, rm, rv, None)
    gathered = torch.stack([ops.bn_stats_local(p, None) for p in parts])
    rm2, rv2 = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    mean2, rstd2, inv_n = ops.bn_sync_finalize(gathered, 1e-5, 0.1, rm2, rv2)
    close(mean2, mean, 1e-6), close(rstd2, rstd, 1e-5), close(rm2, rm, 1e-6), close(rv2, rv, 1e-5)
    assert abs(float(inv_n) * whole.shape[0] - 1.0) < 1e-6
    want = torch.nn.functional.batch_norm(whole, None, None, training=True)
    got = (whole - mean2) * rstd2
    close(got, want, 1e-4)


def _model_and_batches(conv, n_batches=3, graphs=12, hidden=32, layers=2, seed=0):
    from examples.zinc_models import SpModel
    from pygho_b200.hodata.device import attach_host_plans, sp_datadict
    from pygho_b200.hodata.synthetic import make_batch
    from pygho_b200.honn.SpOperator import parse_precomputekey
    torch.manual_seed(seed)
    model = SpModel(conv, num_layer=layers, hiddim=hidden).to(DEV)
    keys = parse_precomputekey(model)
    hbs = [make_batch(graphs, seed=40 + i) for i in range(n_batches)]
    for hb in hbs:
        attach_host_plans(hb, sp_datadict(hb, DEV, keys), keys)
    return model, keys, hbs


@pytest.mark.parametrize("conv,hidden", [("SSWL", 128), ("SSWL", 32), ("DSSGNN", 32), ("NGNN", 32),
                                         ("PPGN", 32)])
def test_padded_batch_equals_exact_batch(conv, hidden):
    """One training step on a capacity-padded batch (inert pads, device row counts) gives the
    loss, predictions and parameter gradients of the step on the exact-size batch."""
    from pygho_b200 import static as ST
    from pygho_b200.hodata.device import sp_datadict
    model, keys, hbs = _model_and_batches(conv, hidden=hidden)
    caps = ST.capacities(hbs, keys, margin=0.05)
    ref = copy.deepcopy(model)
    for hb in hbs[:2]:
        dd = sp_datadict(hb, DEV, keys)
        pred = ref(dd)
        loss = torch.nn.functional.l1_loss(dd["y"].unsqueeze(-1), pred)
        ref.zero_grad()
        loss.backward()
        pdd = ST.attach_registry(sp_datadict(ST.pad_host_batch(hb, caps, keys), DEV, keys), caps)
        B = pdd["num_valid_graphs"]
        with ST.static_shapes(pdd):
            ppred = model(pdd)
            ploss = torch.nn.functional.l1_loss(pdd["y"].unsqueeze(-1)[:B], ppred[:B])
            model.zero_grad()
            ploss.backward()
        assert ppred.shape[0] == B + 1
        close(ppred[:B], pred, 2e-5)
        close(ploss, loss, 2e-5)
        for (k, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
            if q.grad is None:
                assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
                continue
            close(p.grad, q.grad, 2e-4)
        for (k, p), (_, q) in zip(model.named_buffers(), ref.named_buffers()):
            close(p.float(), q.float(), 1e-5)


@pytest.mark.parametrize("threaded", [False, True])
def test_static_feeder_replays_match_eager_exact_steps(threaded):
    """Host-fed training through StaticFeeder (graph replay on padded batches, side-stream
    loader) follows the same trajectory as eager steps on the exact-size batches."""
    from pygho_b200 import static as ST
    from pygho_b200.dist import FlatGradBucket
    from pygho_b200.hodata.device import sp_datadict
    model, keys, hbs = _model_and_batches("SSWL", n_batches=4, graphs=10, hidden=128, layers=2)
    ref = copy.deepcopy(model)

    def make_step(m):
        bucket = FlatGradBucket(m.parameters())
        opt = torch.optim.AdamW(m.parameters(), lr=1e-3, fused=True, capturable=True)

        def step(dd):
            bucket.zero()
            pred, y = m(dd), dd["y"].unsqueeze(-1)
            nv = dd.get("num_valid_graphs")
            if nv is not None:
                pred, y = pred[:nv], y[:nv]
            loss = torch.nn.functional.l1_loss(y, pred)
            loss.backward()
            opt.step()
            return loss.detach()
        return step

    caps = ST.capacities(hbs, keys, margin=0.02)
    padded = [ST.pad_host_batch(hb, caps, keys) for hb in hbs]
    # the feeder warms up and captures on batches 0 and 1 (2 eager warm-up steps per slot):
    # give the reference the same 4 updates first
    ref_step = make_step(ref)
    for i in (0, 0, 1, 1):
        ref_step(sp_datadict(hbs[i], DEV, keys))
    feeder = ST.StaticFeeder(padded, caps, torch.device(DEV, 0), keys, make_step(model), threaded=threaded)
    assert feeder.launches > 10
    got, want = [], []
    for i in range(9):
        got.append(float(feeder.step()))
        want.append(float(ref_step(sp_datadict(hbs[i % len(hbs)], DEV, keys))))
    feeder.close()
    for a, b in zip(got[:3], want[:3]):                  # before Adam amplifies rounding noise
        assert abs(a - b) <= 2e-5 * max(1.0, abs(b)), (got, want)
    for a, b in zip(got, want):
        assert abs(a - b) <= 2e-3 * max(1.0, abs(b)), (got, want)


_SYNCBN_WORKER = r"""
import os, sys, copy
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["PGH_ROOT"])
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
ngpu = torch.cuda.device_count()
dev = torch.device("cuda", rank % ngpu)
torch.cuda.set_device(dev)
backend = "nccl" if ngpu >= world else "gloo"
dist.init_process_group(backend, rank=rank, world_size=world)
from examples.zinc_models import SpModel
from pygho_b200.dist import FlatGradBucket, broadcast_parameters, enable_sync_batchnorm
from pygho_b200.hodata.device import sp_datadict
from pygho_b200.hodata.synthetic import collate, make_graphs
from pygho_b200.honn.SpOperator import parse_precomputekey
torch.manual_seed(0)
model = SpModel("SSWL", num_layer=2, hiddim=32).to(dev)
broadcast_parameters(model)
keys = parse_precomputekey(model)
graphs = make_graphs(16, seed=77)
whole = copy.deepcopy(model)
# single-process reference on the whole batch
dd = sp_datadict(collate(graphs), dev, keys)
loss_w = torch.nn.functional.l1_loss(dd["y"].unsqueeze(-1), whole(dd))
loss_w.backward()
# sharded: rank r takes graphs [r*B/W, (r+1)*B/W), SyncBN, gradient all-reduce (mean)
assert enable_sync_batchnorm(model) > 0
per = len(graphs) // world
sd = sp_datadict(collate(graphs[rank * per:(rank + 1) * per]), dev, keys)
bucket = FlatGradBucket(model.parameters())
bucket.zero()
loss = torch.nn.functional.l1_loss(sd["y"].unsqueeze(-1), model(sd))
loss.backward()
bucket.allreduce_mean()
lt = loss.detach().clone()
dist.all_reduce(lt)
lt /= world
ok = abs(float(lt) - float(loss_w)) <= 1e-4 * max(1.0, abs(float(loss_w)))
worst = 0.0
for (k, p), (_, q) in zip(model.named_parameters(), whole.named_parameters()):
    if q.grad is None:
        continue
    scale = max(1.0, float(q.grad.abs().max()))
    worst = max(worst, float((p.grad - q.grad).abs().max()) / scale)
for (k, p), (_, q) in zip(model.named_buffers(), whole.named_buffers()):
    if p.dtype.is_floating_point:
        worst = max(worst, float((p - q).abs().max()) / max(1.0, float(q.abs().max())))
# replicas stay bit-identical
hi, lo = bucket.flat.clone(), bucket.flat.clone()
dist.all_reduce(hi, op=dist.ReduceOp.MAX)
dist.all_reduce(lo, op=dist.ReduceOp.MIN)
same = bool(torch.equal(hi, lo))
print(f"RESULT rank={rank} backend={backend} ok={ok} worst={worst:.3e} same={same}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if (ok and worst <= 1e-4 and same) else 1)
"""


def test_syncbn_sharded_training_equals_single_process(tmp_path):
    """2 ranks (NCCL on 2 GPUs when the box has them, else gloo moving CUDA tensors of one GPU):
    graph-sharded training with SyncBN + gradient all-reduce == single-process training on the
    whole batch within 1e-4 (loss, every gradient, running statistics); replicas bit-identical."""
    script = tmp_path / "syncbn_worker.py"
    script.write_text(_SYNCBN_WORKER)
    port = 29500 + (os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), PGH_ROOT=ROOT, OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, cwd=ROOT,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    assert all(p.returncode == 0 for p in procs), "\n".join(o[-3000:] for o in outs)


@pytest.mark.parametrize("world", [1, 4])
def test_flat_adamw_equals_torch_adamw(world):
    """FlatAdamW (one launch over flat parameter / gradient / moment buffers, the 1 / world of the
    gradient average folded in) follows torch.optim.AdamW step for step."""
    import copy
    from pygho_b200.dist import FlatAdamW, FlatGradBucket
    torch.manual_seed(5)
    a = torch.nn.Sequential(torch.nn.Linear(37, 53), torch.nn.SiLU(), torch.nn.Linear(53, 3)).to(DEV)
    b = copy.deepcopy(a)
    bucket = FlatGradBucket(a.parameters())
    flat = FlatAdamW(bucket, lr=3e-3, weight_decay=0.05)
    ref = torch.optim.AdamW(b.parameters(), lr=3e-3, weight_decay=0.05)
    for p, q in zip(a.parameters(), b.parameters()):
        assert torch.equal(p, q)                     # re-pointing the parameters kept their values
        assert p.data_ptr() % 16 == 0
    for it in range(6):
        x = torch.randn(64, 37, device=DEV)
        bucket.zero()
        (a(x).square().sum() * world).backward()     # "sum over ranks" of identical shards
        flat.step(1.0 / world)
        ref.zero_grad()
        b(x).square().sum().backward()
        ref.step()
        for p, q in zip(a.parameters(), b.parameters()):
            assert torch.allclose(p, q, rtol=2e-5, atol=2e-6), (it, float((p - q).abs().max()))
    assert float(flat.steps) == 6.0
