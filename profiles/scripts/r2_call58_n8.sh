#!/usr/bin/env bash
# Round 2, GPU call 58 (8 GPUs): N = 8 and N = 1 lines with the final loader (v5).
set -x
O=gpurun_out; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 \
  bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-stock-gpu > $O/r2c58_bench_n8.json 2> $O/r2c58_bench_n8.err
timeout 300 python bench.py --no-cpu-baseline --no-stock-gpu > $O/r2c58_bench_n1.json 2> $O/r2c58_bench_n1.err
python - <<'PY'
import json
for n in (8, 1):
    d=json.loads(open(f"gpurun_out/r2c58_bench_n{n}.json").read().strip().splitlines()[-1]); e=d["e2e"]
    print(n, "value", round(d["value"]), round(d["ms_per_step"],3), "e2e", round(e["value"]), round(e["ms_per_step"],3), e["step_ms"], d["clocks"])
PY
