"""Host-side cost of one SSWL+ training step vs. its device time, and whether the whole step
(forward, loss, backward, fused AdamW) survives CUDA-graph capture.

    python profiles/graph_probe.py            # prints one JSON object

`host_ms` is the wall time the Python thread needs to ENQUEUE a step (no synchronisation
inside the loop); `device_ms` the CUDA-event time of the same steps.  If host_ms is close to
device_ms the step is launch-bound as soon as several ranks share the host's cores, which is
what graph replay removes."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from examples.zinc_models import SpModel  # noqa: E402
from pygho_b200.dist import FlatGradBucket  # noqa: E402
from pygho_b200.hodata.device import prefetch_plans, sp_datadict  # noqa: E402
from pygho_b200.hodata.synthetic import make_batch  # noqa: E402
from pygho_b200.honn.SpOperator import parse_precomputekey  # noqa: E402

B = int(os.environ.get("BATCH", "1024"))
STEPS = int(os.environ.get("STEPS", "20"))
dev = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = True
torch.manual_seed(0)
model = SpModel("SSWL", num_layer=6, hiddim=128).to(dev)
keys = parse_precomputekey(model)
bucket = FlatGradBucket(model.parameters())
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=True, capturable=True)
dd = sp_datadict(make_batch(B, seed=0), dev, keys)
prefetch_plans(dd, keys)


def step():
    bucket.zero()
    loss = torch.nn.functional.l1_loss(dd["y"].unsqueeze(-1), model(dd))
    loss.backward()
    opt.step()
    return loss


def measure(fn, steps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    host = (time.perf_counter() - t0) * 1e3 / steps
    torch.cuda.synchronize()
    return host, e0.elapsed_time(e1) / steps


out = {"batch": B, "steps": STEPS, "cores": len(os.sched_getaffinity(0))}
for _ in range(5):
    step()
out["eager_host_ms"], out["eager_device_ms"] = measure(step, STEPS)

try:
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        static_loss = step()
    graph.replay()
    torch.cuda.synchronize()
    l0 = float(static_loss)
    out["graph_host_ms"], out["graph_device_ms"] = measure(graph.replay, STEPS)
    l1 = float(static_loss)
    out["graph_loss_first"], out["graph_loss_last"] = l0, l1
    out["graph_ok"] = bool(l1 == l1 and l1 < l0 * 1.5)
except Exception as e:  # noqa: BLE001
    out["graph_ok"] = False
    out["graph_error"] = f"{type(e).__name__}: {e}"[:400]
print(json.dumps(out))
