"""GPU tests of the TMA / tcgen05 Linear layer with the BatchNorm statistics in its epilogue
(csrc/linear_stats.cu): the GEMM against torch (TF32 tolerance, stated), the statistics against
the kernel's own output (they must describe y exactly), pad rows, SyncBN triples, and the MLP
block built on it against the cuBLAS + separate-statistics path."""
import copy

import numpy as np
import pytest
import torch

import pygho_b200.ops  # noqa: F401  (registers torch.ops.pygho_b200.*)

pytestmark = pytest.mark.gpu
DEV = "cuda"
TF32_TOL = 2e-3          # relative to the largest |y|: 10-bit mantissa products, fp32 accumulation


def close(a, b, rtol):
    a, b = a.detach().double().cpu().numpy(), b.detach().double().cpu().numpy()
    assert a.shape == b.shape, (a.shape, b.shape)
    scale = max(1.0, float(np.abs(b).max()) if b.size else 1.0)
    err = float(np.abs(a - b).max(initial=0.0))
    assert err <= rtol * scale, (err, scale)


@pytest.fixture(autouse=True)
def _tf32():
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    yield
    torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("M,K,N", [(1, 32, 128), (127, 128, 128), (128, 128, 128), (129, 384, 128),
                                   (4097, 128, 128), (70001, 384, 128), (300000, 128, 128),
                                   (20000, 64, 128), (70001, 384, 384), (4097, 384, 384),
                                   (33333, 256, 256), (257, 128, 384), (9000, 256, 128)])
@pytest.mark.parametrize("has_bias", [True, False])
def test_linear_stats_matches_torch_and_its_own_output(M, K, N, has_bias):
    ops = torch.ops.pygho_b200
    g = torch.Generator(device=DEV).manual_seed(M + K + N)
    x = torch.randn(M, K, device=DEV, generator=g) * 1.5 + 0.3
    w = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5
    b = torch.randn(N, device=DEV, generator=g) if has_bias else None
    rm, rv = torch.zeros(N, device=DEV), torch.ones(N, device=DEV)
    nbt = torch.zeros((), dtype=torch.int64, device=DEV)
    y, mean, rstd = ops.linear_stats(x, w, b, 1e-5, 0.1, rm, rv, None, nbt, False)
    want = torch.nn.functional.linear(x.double(), w.double(), None if b is None else b.double())
    close(y, want, TF32_TOL)
    # statistics describe the y that was written (independent of the GEMM precision)
    yd = y.double()
    close(mean, yd.mean(0), 1e-5)
    var = yd.var(0, unbiased=False)
    # the epilogue accumulates sums of (y - bias) and of its squares per 32-row slice in fp32
    # (then in double): the variance carries a rounding error of ~1e-7 * E[(y - bias)^2]
    got_var = 1.0 / rstd.double() ** 2 - 1e-5
    bd = 0.0 if b is None else b.double()
    bound = 4e-7 * ((yd - bd) ** 2).mean(0) + 1e-9
    assert bool(((got_var - var).abs() <= bound).all()), float(((got_var - var).abs() / bound).max())
    close(rm, 0.1 * yd.mean(0), 1e-5)
    if M > 1:
        close(rstd, 1.0 / torch.sqrt(var + 1e-5), 2e-5)
        close(rv, 0.9 + 0.1 * yd.var(0, unbiased=True), 2e-5)
    assert int(nbt) == 1
    # deterministic
    y2, mean2, rstd2 = ops.linear_stats(x, w, b, 1e-5, 0.1, None, None, None, None, False)
    assert torch.equal(y, y2) and torch.equal(mean, mean2) and torch.equal(rstd, rstd2)


def test_linear_stats_pad_rows_and_syncbn_triples():
    ops = torch.ops.pygho_b200
    g = torch.Generator(device=DEV).manual_seed(9)
    M, cap, K = 5000, 5200, 384
    x = torch.randn(cap, K, device=DEV, generator=g)
    x[M:] *= 50.0                                      # junk in the pad rows
    w = torch.randn(128, K, device=DEV, generator=g) / K ** 0.5
    b = torch.randn(128, device=DEV, generator=g)
    n = torch.tensor([M], dtype=torch.int32, device=DEV)
    y, mean, rstd = ops.linear_stats(x[:M].contiguous(), w, b, 1e-5, 0.1, None, None, None, None, False)
    yp, meanp, rstdp = ops.linear_stats(x, w, b, 1e-5, 0.1, None, None, n, None, False)
    assert torch.equal(yp[:M], y)
    close(meanp, mean, 1e-6), close(rstdp, rstd, 1e-6)
    # rank-local triples (mean, M2, count) merged by bn_sync_finalize == whole-batch statistics
    parts = [x[:1700].contiguous(), x[1700:M].contiguous()]
    trip = []
    for p in parts:
        yy, local, _e = ops.linear_stats(p, w, b, 1e-5, 0.1, None, None, None, None, True)
        assert local.shape == (3, 128) and float(local[2, 0]) == p.shape[0]
        trip.append(local)
    m2, r2, inv_n = ops.bn_sync_finalize(torch.stack(trip), 1e-5, 0.1, None, None)
    close(m2, mean, 1e-6), close(r2, rstd, 1e-5)
    assert abs(float(inv_n) * M - 1.0) < 1e-6


@pytest.mark.parametrize("rows,cin", [(6000, 128), (33333, 384), (20000, 256)])
def test_mlp_block_on_fused_gemm_equals_cublas_path(rows, cin):
    """MLP (Linear -> BN -> SiLU) x 2 with the fused GEMM + statistics against the same block on
    cuBLAS TF32 + separate statistics pass: outputs and all gradients within TF32 tolerance."""
    from pygho_b200 import ops
    from pygho_b200.honn.utils import MLP
    torch.manual_seed(rows)
    mlp = MLP(cin, 128, 2, True, norm="bn", act="silu", normparam=0.3).to(DEV)
    # hidden block cin -> cin and output block cin -> 128: both on the fused GEMM
    ref = copy.deepcopy(mlp)
    x = torch.randn(rows, cin, device=DEV)
    w = torch.randn(rows, 128, device=DEV)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    last = [m for m in mlp.modules() if isinstance(m, torch.nn.Linear)][-1]
    wide, ops._FUSED_GEMM_WIDE = ops._FUSED_GEMM_WIDE, True    # cover N = 256 / 384 and K = 128 too
    try:
        assert ops.linear_stats_ok(torch.empty(rows, last.in_features, device=DEV), last.weight)
        (mlp(xa) * w).sum().backward()
    finally:
        ops._FUSED_GEMM_WIDE = wide
    ops.set_fused_linear_stats(False)
    try:
        (ref(xb) * w).sum().backward()
    finally:
        ops.set_fused_linear_stats(True)
    close(xa.grad, xb.grad, 5e-3)
    for (k, p), (_, q) in zip(mlp.named_parameters(), ref.named_parameters()):
        if k.endswith("bias") and "norm" not in k:
            continue
        close(p.grad, q.grad, 5e-3)
    for (k, p), (_, q) in zip(mlp.named_buffers(), ref.named_buffers()):
        close(p.float(), q.float(), 2e-3)
