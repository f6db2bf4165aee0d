#!/usr/bin/env bash
# Round 2, GPU call 13: forked launches for independent kernels; 128-graph and 1024-graph steps.
set -x
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c13_tests.log 2>&1; echo "rc=$?" >> $O/r2c13_tests.log
tail -8 $O/r2c13_tests.log
for fork in 1 0; do
  PYGHO_B200_FORK=$fork timeout 600 python bench.py --no-cpu-baseline --no-stock-gpu --no-roofline > $O/r2c13_bench_sswl_fork$fork.json 2> $O/r2c13_bench_sswl_fork$fork.err
  PYGHO_B200_FORK=$fork timeout 600 python bench.py --batch 128 --no-cpu-baseline --no-stock-gpu --no-roofline > $O/r2c13_bench_sswl128_fork$fork.json 2> $O/r2c13_bench_sswl128_fork$fork.err
done
python - <<'PY'
import json
for f in ("sswl_fork1","sswl_fork0","sswl128_fork1","sswl128_fork0"):
    try:
        d=json.loads(open(f"gpurun_out/r2c13_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d["e2e"]["step_ms"], d["clocks"])
    except Exception as e: print(f, "ERR", e)
PY
