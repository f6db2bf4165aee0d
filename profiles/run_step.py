"""One SSWL+ training step (B=1024 graphs) bracketed by cudaProfilerStart/Stop, for
`ncu --profile-from-start off` (launch list of exactly one step) -- see profiles/README.md."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from examples.zinc_models import SpModel  # noqa: E402
from pygho_b200.dist import FlatGradBucket  # noqa: E402
from pygho_b200.hodata.device import sp_datadict  # noqa: E402
from pygho_b200.hodata.synthetic import make_batch  # noqa: E402
from pygho_b200.honn.SpOperator import parse_precomputekey  # noqa: E402

B = int(os.environ.get("BATCH", "1024"))
dev = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = True
torch.manual_seed(0)
model = SpModel("SSWL", num_layer=6, hiddim=128).to(dev)
keys = parse_precomputekey(model)
bucket = FlatGradBucket(model.parameters())
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=True)
dd = sp_datadict(make_batch(B, seed=0), dev, keys)


def step():
    bucket.zero()
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
