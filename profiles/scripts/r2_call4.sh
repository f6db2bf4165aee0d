#!/usr/bin/env bash
# Round 2, GPU call 4: faster ticket tail, slab-sum kernel, cycle-free loader; traces.
set -x
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r2c4_tests.log 2>&1; echo "rc=$?" >> $O/r2c4_tests.log
tail -6 $O/r2c4_tests.log
export PYGHO_B200_BENCH_TRACE=1
python bench.py --no-cpu-baseline --no-stock-gpu > $O/r2c4_bench_sswl.json 2> $O/r2c4_bench_sswl.err; grep trace $O/r2c4_bench_sswl.err
python bench.py --batch 128 --no-cpu-baseline --no-stock-gpu > $O/r2c4_bench_sswl128.json 2> $O/r2c4_bench_sswl128.err; grep trace $O/r2c4_bench_sswl128.err
BATCH=128 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \
  --csv --log-file $O/r2c4_step128_launches.csv python profiles/run_step.py > $O/r2c4_step128.log 2>&1
python - <<'PY'
import json
for f in ("r2c4_bench_sswl","r2c4_bench_sswl128"):
    d=json.load(open(f"gpurun_out/{f}.json"))
    print(f, round(d["value"]), round(d["ms_per_step"],3), d["gpu_launches"]/d["steps"], "e2e", round(d["e2e"]["value"]), d["e2e"]["step_ms"])
PY
python profiles/launch_summary.py $O/r2c4_step128_launches.csv 16
BATCH=1024 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \
  --csv --log-file $O/r2c4_step1024_launches.csv python profiles/run_step.py > $O/r2c4_step1024.log 2>&1
python profiles/launch_summary.py $O/r2c4_step1024_launches.csv 16
