"""Train the SSWL+ model of the reference's ``example/zinc.py`` on synthetic ZINC-shaped batches
with pygho_b200: host batches -> ``DevicePrefetcher`` (worker thread + side stream) -> eager
steps, or, with ``--resident``, one CUDA-graph replay per step on device-resident batches.

    python examples/train_zinc.py --steps 50 [--resident] [--conv SSWL|NGNN|DSSGNN] [--batch 1024]
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from examples.zinc_models import SpModel  # noqa: E402
from pygho_b200.dist import FlatGradBucket  # noqa: E402
from pygho_b200.graph import StepGraph  # noqa: E402
from pygho_b200.hodata.device import (DeferredScalar, DevicePrefetcher, attach_host_plans,  # noqa: E402
                                      prefetch_plans, sp_datadict)
from pygho_b200.hodata.synthetic import make_batch  # noqa: E402
from pygho_b200.honn.SpOperator import parse_precomputekey  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--conv", default="SSWL")
    ap.add_argument("--hidden", type=int, default=128)
    ap.add_argument("--layers", type=int, default=6)
    ap.add_argument("--num-batches", type=int, default=4)
    ap.add_argument("--resident", action="store_true", help="batches stay in HBM, steps replay CUDA graphs")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.backends.cuda.matmul.allow_tf32 = True          # like the reference (zinc.py:30)
    torch.manual_seed(0)
    model = SpModel(args.conv, num_layer=args.layers, hiddim=args.hidden).to(dev)
    keys = parse_precomputekey(model)
    tables = {"x": model.x_encoder.num_embeddings, "A": model.ea_encoder.num_embeddings,
              "X": model.tuplefeat_encoder.num_embeddings}
    bucket = FlatGradBucket(model.parameters())
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=True, capturable=True)
    hbs = [make_batch(args.batch, seed=i) for i in range(args.num_batches)]
    print(f"{args.conv}: {sum(p.numel() for p in model.parameters())} parameters, "
          f"{len(hbs)} batches of {args.batch} graphs ({hbs[0].tupleid.shape[1]} tuples)")

    def train_step(dd):
        bucket.zero()
        loss = torch.nn.functional.l1_loss(dd["y"].unsqueeze(-1), model(dd))
        loss.backward()
        opt.step()
        return loss.detach()

    # plans are built on the device the first time a batch is seen and shipped from the host
    # afterwards, like the reference's pre-transformed datasets
    dds = []
    for hb in hbs:
        dd = sp_datadict(hb, dev, keys)
        attach_host_plans(hb, dd, keys)
        prefetch_plans(dd, keys, embeddings=tables)
        dds.append(dd)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if args.resident:
        graphs = [StepGraph(lambda dd=dd: train_step(dd), warmup=1) for dd in dds]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(args.steps):
            loss = graphs[i % len(graphs)].replay()
            if i % 10 == 0:
                print(f"step {i:4d} loss {float(loss):.4f}")
    else:
        feeder = DevicePrefetcher(hbs, dev, keys, embeddings=tables)
        reader = DeferredScalar()
        for i in range(args.steps):
            dd = feeder.get()
            loss = train_step(dd)
            prev = reader.push(loss)            # loss of step i-1; waits for that step only
            if prev is not None and (i - 1) % 10 == 0:
                print(f"step {i - 1:4d} loss {prev:.4f}")
            feeder.advance()                    # H2D + plans of the next batch, side stream
        print(f"step {args.steps - 1:4d} loss {reader.flush():.4f}")
        feeder.close()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{args.steps} steps in {dt:.2f} s = {args.batch * args.steps / dt:.0f} graphs/s")


if __name__ == "__main__":
    main()
