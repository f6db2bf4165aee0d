"""Debug probe for the TMA-staged mamamm (algo 3): structured tiny inputs whose outputs reveal
which operand element reached which accumulator position."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import pygho_b200.ops  # noqa
dev = "cuda"
torch.set_printoptions(linewidth=200, precision=1, sci_mode=False)
op = torch.ops.pygho_b200.mamamm
for (n, d) in ((8, 4), (8, 8), (16, 4)):
    mask = torch.ones(1, n, n, dtype=torch.bool, device=dev)
    A = torch.ones(1, n, n, d, device=dev)
    B = torch.ones(1, n, n, d, device=dev)
    print(f"n={n} d={d} all-ones: expect {n}; got unique", torch.unique(op(A, False, B, False, mask, None, 3)).tolist()[:10])
    A = torch.zeros(1, n, n, d, device=dev); B = torch.zeros(1, n, n, d, device=dev)
    ar = torch.arange(n, device=dev, dtype=torch.float32)
    ch = torch.arange(d, device=dev, dtype=torch.float32)
    A[0, :, 0, :] = (ar[:, None] + 1) * 10        # A[i, j=0, c] = 10 (i+1)
    B[0, 0, :, :] = (ar[:, None] + 1) + 0.1 * (ch[None, :] + 1)   # B[j=0, k, c] = (k+1) + 0.1 (c+1)
    out = op(A, False, B, False, mask, None, 3)
    ref = op(A, False, B, False, mask, None, 0)
    print("max err vs fp32:", float((out - ref).abs().max()), "ref max", float(ref.abs().max()))
    print("out[0,:,:,0]\n", out[0, :, :, 0]); print("ref[0,:,:,0]\n", ref[0, :, :, 0])
    print("out[0,1,:, :]\n", out[0, 1, :, :]); print("ref[0,1,:,:]\n", ref[0, 1, :, :])
    # j-dependence: A[i, j, c] = 1 only at j = 3
    A.zero_(); B.zero_()
    A[0, :, 3, :] = 1.0
    B[0, :, :, :] = (ar[:, None, None] + 1) * 100 + (ar[None, :, None] + 1)    # B[j,k,c] = 100 (j+1) + (k+1)
    out = op(A, False, B, False, mask, None, 3); ref = op(A, False, B, False, mask, None, 0)
    print("j-probe out[0,0,:,0]", out[0, 0, :, 0].tolist(), "ref", ref[0, 0, :, 0].tolist())
torch.cuda.synchronize()
