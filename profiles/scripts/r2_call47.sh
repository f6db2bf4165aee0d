#!/usr/bin/env bash
# Round 2, GPU call 47: final bench lines of the round, all four workloads (v3), N = 1.
set -x
O=gpurun_out; mkdir -p $O
for W in sswl ppgn_dd dssgnn_sr25 i2_sr25; do
  timeout 900 python bench.py --workload $W > $O/r2c47_bench_${W}.json 2> $O/r2c47_bench_${W}.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/r2c47_bench_${W}.json").read().strip().splitlines()[-1])
    r=d.get("roofline") or {}; s=d.get("stock_gpu_baseline") or {}; c=d.get("cpu_baseline") or {}
    print("$W", "value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "roofline", r.get("kernel","")[:40], round(r.get("frac",0),3), r.get("us_per_launch"), "stock", s.get("own_over_stock"), "cpu", c.get("value"), c.get("cores"))
except Exception as ex: print("$W ERR", ex)
PY
done
