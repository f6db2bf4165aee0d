"""One dense PPGN (MaModel, bench.py --workload ppgn_dd) training step bracketed by
cudaProfilerStart/Stop, for `ncu --profile-from-start off` (launch list of exactly one step)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from examples.zinc_models import MaModel  # noqa: E402
from pygho_b200.hodata.device import ma_datadict  # noqa: E402
from pygho_b200.hodata.synthetic import make_batch  # noqa: E402

B = int(os.environ.get("BATCH", "128"))
dev = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = True
torch.manual_seed(0)
model = MaModel("PPGN", num_layer=6, hiddim=128).to(dev)
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=True)
dd = ma_datadict(make_batch(B, seed=0), dev)


def step():
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.l1_loss(dd["y"].unsqueeze(-1), model(dd))
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
