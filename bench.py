#!/usr/bin/env python
"""Benchmark: SSWL+ training throughput (graphs/s) on synthetic ZINC-shaped batches.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA kernels)
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU path (oracle port)

One "step" is one full optimisation step (forward, L1 loss, backward, gradient
all-reduce when N > 1, AdamW) of the 6-layer SSWL+ sparse model of the reference's
``example/zinc.py`` (hidden 128, mlplayer 2, outlayer 4, bn/silu, npool sum, lpool mean)
on one batch of ``--batch`` graphs per GPU.  Prints ONE JSON line (rank 0).

* ``value``: graphs/s with batches and plans resident in HBM (device-timed, max over ranks).
* ``e2e``: the same step driven from pinned HOST buffers in the reference's datadict
  format (x, edge_index, edge_attr, tupleid, tuplefeat, batch, y and the precomputed
  ``acd`` plans): host->device copies, SparseTensor wrapping, CSR regrouping of the plans
  and a device->host read of the loss are inside the timed region.
* ``roofline``: the dominant kernel of this path, the fused spspmm forward
  (``pgh_seg_gmr_f32``), timed alone with CUDA events on rotating operand sets larger
  than L2; achieved = algorithmic bytes / mean launch duration.
* ``cpu_baseline``: the oracle port of the reference (torch CPU ops) on the host cores,
  on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# batches differ in size from step to step: growable segments keep the caching allocator from
# falling back to cudaMalloc/cudaFree (device-synchronising) when a block does not fit, and
# size classes of 1/8 power of two let batches of slightly different sizes reuse each other's
# blocks instead of fragmenting the pool (every growth is a cuMemMap stall of 10-100 ms,
# profiles/e2e_trace.py)
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF",
                      "expandable_segments:True,roundup_power2_divisions:8")

# NCCL prints its version banner (NCCL_DEBUG=VERSION/INFO) to stdout by default: keep stdout
# for the one JSON line
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import numpy as np  # noqa: E402
import torch  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sswl_plus_zinc_shape_train_graphs_per_s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="graphs per GPU per step")
    ap.add_argument("--conv", default="SSWL")
    ap.add_argument("--hidden", type=int, default=128)
    ap.add_argument("--layers", type=int, default=6)
    ap.add_argument("--num-batches", type=int, default=3, help="distinct batches rotated")
    ap.add_argument("--ref-batch", type=int, default=128, help="graphs per CPU reference step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-graph", action="store_true",
                    help="launch the resident-batch steps eagerly instead of replaying CUDA graphs")
    return ap.parse_args()


def workload_name(args, world):
    return (f"{args.conv}+ SpModel {args.layers}x{args.hidden} (zinc.py work.sh flags), synthetic "
            f"ZINC-shape k-hop(3) batches, {args.batch} graphs/GPU x {world} GPU")


# ------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.lines, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, smax, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = max(smax, float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------- reference arm
def oracle_train_throughput(args, steps, warmup, batch_graphs, seed=12345):
    """graphs/s of the CPU restatement of the reference (oracle/model_oracle.py)."""
    from oracle import model_oracle as MO
    from oracle import pygho_oracle as O
    from pygho_b200.hodata.synthetic import make_batch
    # all host threads the process may use (torchrun pins OMP_NUM_THREADS=1 by default)
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except Exception:
        pass
    torch.manual_seed(0)
    hb = make_batch(batch_graphs, seed=seed)
    keys = {"SSWL": ["X___X___1___A___0", "X___A___1___X___0"], "NGNN": ["X___X___1___A___0"],
            "DSSGNN": ["X___X___1___A___0"], "PPGN": ["X___X___1___X___0"]}[args.conv]
    plans = {}
    for key in keys:
        _o0, o1, d1, o2, d2 = key.split("___")
        pick = lambda op: hb.edge_index if op == "A" else hb.tupleid  # noqa: E731
        plans[key + "___acd"] = torch.from_numpy(
            O.filterind(hb.tupleid, *O.spspmm_ind(pick(o1), int(d1), pick(o2), int(d2))))
    g = MO.host_graph_dict(hb, plans)
    model = MO.OSpModel(conv=args.conv, num_layer=args.layers, hiddim=args.hidden)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)

    def step():
        opt.zero_grad()
        loss = torch.nn.functional.l1_loss(g["y"].unsqueeze(-1), model(g))
        loss.backward()
        opt.step()
        return float(loss.detach())

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch_graphs * steps / dt, dt / steps * 1e3, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    # bound the whole run to a few minutes: the CPU path does ~100-400 graphs/s
    budget_steps = max(2, min(steps, 24))
    val, ms, cores = oracle_train_throughput(args, budget_steps, min(warmup, 3), args.ref_batch)
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    sample = (f"{budget_steps} steps of {args.ref_batch} graphs (same generator and model as "
              f"the GPU arm, which runs {args.batch} graphs/GPU)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "graphs/s",
        "n_gpus": args.gpus, "steps": budget_steps, "warmup": min(warmup, 3), "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args, world), "l2": "cpu run"},
        "cpu_baseline": {"value": val, "unit": "graphs/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": val, "unit": "graphs/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }), flush=True)


# ------------------------------------------------------------------------ own arm
def roofline_spspmm(dd_list, hidden, device, peaks):
    """Time the fused spspmm forward (the SSWL `X x A` key) alone: CUDA events around
    `iters` back-to-back launches cycling over operand sets whose total size exceeds L2."""
    from pygho_b200 import plans as P
    ops = torch.ops.pygho_b200
    dd = dd_list[0]
    acd = dd["X___X___1___A___0___acd"]
    nX, nA = dd["X"].nnz, dd["A"].nnz
    plan = P.plan_from_acd(acd, nX, nX, nA)
    g = plan.group("a")
    T = plan.T
    alg_bytes = 4 * hidden * (nX + nA + nX) + 4 * (2 * T + nX + 1)
    l2 = 128 << 20
    try:
        from pygho_b200 import _lib
        import ctypes
        info = (ctypes.c_int32 * 5)()
        _lib.call("pgh_device_info", info)
        l2 = int(info[2]) << 10
    except Exception:
        pass
    per_set = 4 * hidden * (2 * nX + nA)
    nsets = max(4, int(np.ceil(2.5 * l2 / per_set)))
    gen = torch.Generator(device=device).manual_seed(0)
    sets = [(torch.randn((nX, hidden), device=device, generator=gen),
             torch.randn((nA, hidden), device=device, generator=gen)) for _ in range(nsets)]
    for xv, av in sets[:3]:
        ops.seg_gmr(xv, g.first, None, av, g.second, g.rowptr, nX, 0)
    iters = 5 * nsets
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def launches():
        for i in range(iters):
            xv, av = sets[i % nsets]
            ops.seg_gmr(xv, g.first, None, av, g.second, g.rowptr, nX, 0)

    # the `iters` launches are captured into one CUDA graph and replayed, so the events bracket
    # back-to-back kernel executions without Python / launch gaps (eager fallback if capture fails)
    timing = "cuda graph replay of the launches"
    try:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            launches()
        graph.replay()
        torch.cuda.synchronize(device)
        best = float("inf")
        for _ in range(3):
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize(device)
            best = min(best, e0.elapsed_time(e1))
        us = best * 1e3 / iters
        del graph
    except Exception:  # noqa: BLE001
        torch.cuda.synchronize(device)
        timing = "eager launches"
        e0.record()
        launches()
        e1.record()
        torch.cuda.synchronize(device)
        us = e0.elapsed_time(e1) * 1e3 / iters
    achieved = alg_bytes / (us * 1e-6) / 1e9
    peak = peaks.get("hbm_gbs")
    traffic = None
    try:  # dram bytes of this launch from the committed ncu --set full capture
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_roofline_traffic.json")))["traffic"]
    except Exception:
        pass
    return {"bound": "hbm", "kernel": "seg_gmr_lean_kernel<sum,B> (spspmm fwd, key X___X___1___A___0)",
            "achieved": achieved, "peak": peak if peak else 6650.0, "unit": "GB/s",
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst, kernel timed alone)" if peak
            else "fallback 6650 GB/s (B200_PROFILING.md)",
            "frac": achieved / (peak if peak else 6650.0), "traffic": traffic,
            "us_per_launch": us, "algorithmic_bytes": alg_bytes,
            "rows": nX, "triples": T, "operand_sets": nsets, "l2_bytes": l2, "timing": timing}


def run_b200(args):
    import torch.distributed as dist
    from examples.zinc_models import SpModel
    from pygho_b200 import _lib
    from pygho_b200.dist import FlatGradBucket, broadcast_parameters
    from pygho_b200.hodata.device import attach_host_plans, sp_datadict
    from pygho_b200.hodata.synthetic import make_batch
    from pygho_b200.honn.SpOperator import parse_precomputekey

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl b200) needs a CUDA device: the kernels have no CPU path")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    torch.backends.cuda.matmul.allow_tf32 = True      # reference: set_float32_matmul_precision('high')
    torch.backends.cudnn.allow_tf32 = True
    _lib.load()

    torch.manual_seed(0)
    model = SpModel(args.conv, num_layer=args.layers, hiddim=args.hidden).to(device)
    broadcast_parameters(model)
    keys = parse_precomputekey(model)
    bucket = FlatGradBucket(model.parameters())
    # capturable: the step counter lives on the device, so the optimizer can be graph-captured
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=True, capturable=True)

    # synthetic data: `num_batches` distinct batches per rank (weak scaling: fixed per-GPU work)
    hbs = [make_batch(args.batch, seed=1000 * rank + i) for i in range(args.num_batches)]
    pinned = {}
    dds = []
    for hb in hbs:
        dd = sp_datadict(hb, device, keys, pinned)
        attach_host_plans(hb, dd, keys)
        dds.append(dd)
    # pin the host plans too (they are what the reference's loader would ship)
    for hb in hbs:
        for k, v in hb.plans.items():
            pinned[id(v)] = torch.from_numpy(v).pin_memory()
    h2d_bytes = int(np.mean([hb.nbytes() for hb in hbs]))

    def train_step(dd):
        bucket.zero()
        pred = model(dd)
        loss = torch.nn.functional.l1_loss(dd["y"].unsqueeze(-1), pred)
        loss.backward()
        bucket.allreduce_mean()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def timed(fn, steps, finish=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launches()
        e0.record()
        for i in range(steps):
            fn(i)
        if finish is not None:
            finish()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), _lib.launches() - l0

    warm = max(3, args.warmup)
    for i in range(warm):
        train_step(dds[i % len(dds)])
    # Resident batches: the whole step (fwd, loss, bwd, all-reduce, AdamW) of each batch is
    # captured into one CUDA graph and replayed (pygho_b200/graph.py) -- the eager step needs
    # 14.8 ms of host time for 16.4 ms of device time, i.e. it is nearly launch-bound.  If
    # capture fails on this rank the eager step is used (it issues the same collectives).
    graphs, graph_note = None, "eager (--no-graph)"
    if not args.no_graph:
        try:
            from pygho_b200.graph import StepGraph
            graphs = [StepGraph(lambda dd=dd: train_step(dd), warmup=1) for dd in dds]
            graph_note = f"cuda graph replay, one graph per resident batch ({len(graphs)})"
        except Exception as e:  # noqa: BLE001
            graphs = None
            graph_note = f"eager (graph capture failed: {type(e).__name__}: {str(e)[:120]})"
            print(f"[bench] rank {rank}: {graph_note}", file=sys.stderr, flush=True)
            torch.cuda.synchronize(device)
    if graphs is not None:
        run_step = lambda i: graphs[i % len(graphs)].replay()  # noqa: E731
    else:
        run_step = lambda i: train_step(dds[i % len(dds)])  # noqa: E731
    for i in range(warm):
        run_step(i)
    with ClockSampler(local) as clocks:
        ms_total, launches = timed(run_step, args.steps)
    if graphs is not None:                                # replays do not pass through Python
        launches = sum(graphs[i % len(graphs)].launches for i in range(args.steps))
    clock_summary = clocks.summary()
    ms_step = ms_total / args.steps
    value = args.batch * world / (ms_step * 1e-3)

    # ---- end to end: pinned host buffers in, loss out -------------------------------
    # every step's inputs come from pinned host memory (reference datadict format incl. the
    # acd plans); the copies + CSR regrouping of batch i+1 are issued on a side stream right
    # after step i has been launched (DevicePrefetcher), the loss is read back every step
    from pygho_b200.hodata.device import DevicePrefetcher
    tables = {"x": model.x_encoder.num_embeddings, "A": model.ea_encoder.num_embeddings,
              "X": model.tuplefeat_encoder.num_embeddings}
    feeder = DevicePrefetcher(hbs, device, keys, pinned, embeddings=tables)

    # The loss of every step is copied to pinned host memory and read by the host ONE step
    # late (deferred logging): the host launches step i, then waits for step i-1's loss,
    # then issues the prefetch of batch i+1 -- the GPU never idles while the host blocks.
    # That wait is also the synchronisation DevicePrefetcher.advance() asks for (batch i-1
    # is fully consumed before its buffers are recycled).
    from pygho_b200.hodata.device import DeferredScalar
    reader = DeferredScalar()
    losses = []

    def read_pending():
        v = reader.flush()
        if v is not None:
            losses.append(v)                             # D2H read of the result

    def e2e_step(i):
        dd = feeder.get()
        loss = train_step(dd)
        prev = reader.push(loss)                         # loss of the step before this one
        if prev is not None:
            losses.append(prev)
        feeder.advance()                                 # next batch's H2D + plans, side stream

    def e2e_steps(n):
        for i in range(n):
            e2e_step(i)
        read_pending()                                   # every timed step's loss is read

    # two full cycles over the host batches: every allocation size of both streams has been
    # seen (a growth of an expandable segment inside the timed region costs 10-100 ms)
    e2e_steps(max(args.warmup, 2 * len(hbs) + 1))
    if os.environ.get("PYGHO_B200_BENCH_DEBUG"):
        for i in range(12):
            torch.cuda.synchronize(device)
            t0 = time.perf_counter()
            e2e_step(i)
            torch.cuda.synchronize(device)
            print(f"[debug] e2e step {i}: {1e3 * (time.perf_counter() - t0):.2f} ms", file=sys.stderr)
        read_pending()
    losses.clear()
    import gc
    gc.collect()
    gc.freeze()        # long-lived objects out of the collector's way: no multi-ms gen-2 pauses
    trace = [] if os.environ.get("PYGHO_B200_BENCH_TRACE") else None
    timed_step = e2e_step
    if trace is not None:
        def timed_step(i):
            e2e_step(i)
            trace.append(time.perf_counter())
    mem0 = torch.cuda.memory_stats(device)
    e2e_ms, _ = timed(timed_step, args.steps, finish=read_pending)
    if trace is not None:
        mem1 = torch.cuda.memory_stats(device)
        gaps = [round(1e3 * (b - a), 2) for a, b in zip(trace[:-1], trace[1:])]
        grew = {k: mem1[k] - mem0[k] for k in ("num_alloc_retries", "num_device_alloc", "num_device_free")}
        print(f"[trace] e2e host ms between steps: {gaps}; allocator: {grew}", file=sys.stderr)
    assert len(losses) == args.steps and all(np.isfinite(losses)), "e2e: every step's loss is read"
    e2e_value = args.batch * world / (e2e_ms / args.steps * 1e-3)
    feeder.close()                                       # no stray side-stream work below
    torch.cuda.synchronize(device)

    out = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        nX, nA, N = dds[0]["X"].nnz, dds[0]["A"].nnz, hbs[0].num_nodes
        out = {
            "metric": METRIC, "value": value, "unit": "graphs/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(args, world), "global_batch": args.batch * world,
                       "per_gpu_batch": args.batch, "nodes": N, "edges": nA, "tuples": nX,
                       "triples_per_key": int(dds[0][keys[0] + "___acd"].shape[1]),
                       "matmul": "tf32 (reference sets float32_matmul_precision('high'))",
                       "l2": f"step working set >> L2; {len(dds)} distinct batches rotated",
                       "launch": graph_note,
                       "parallelism": f"dp{world} (graphs sharded, flat-bucket NCCL all-reduce)"},
            "e2e": {"value": e2e_value, "unit": "graphs/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches,
            "clocks": clock_summary,
        }
    if rank == 0 and not args.no_roofline:
        out["roofline"] = roofline_spspmm(dds, args.hidden, device, peaks)
    if world > 1:
        dist.barrier()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        del dds
        torch.cuda.empty_cache()
        v, ms, cores = oracle_train_throughput(args, 3, 1, args.ref_batch)
        out["cpu_baseline"] = {"value": v, "unit": "graphs/s", "cores": cores, "kind": "port",
                               "sample": f"3 steps of {args.ref_batch} graphs, oracle port of the "
                                         "reference (torch CPU ops), same model and generator"}
    if rank == 0:
        print(json.dumps(out), flush=True)
    # Orderly teardown: prefetch thread first, then the CUDA graphs (they hold captured NCCL
    # kernels and must go before the communicator), then the process group.  A watchdog ends
    # the process with status 0 if NCCL's shutdown blocks: the result line is already out.
    sys.stdout.flush()
    sys.stderr.flush()
    killer = threading.Timer(30.0, lambda: os._exit(0))
    killer.daemon = True
    killer.start()
    del feeder, run_step
    graphs = None
    import gc
    gc.collect()
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    killer.cancel()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
