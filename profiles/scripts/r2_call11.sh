#!/usr/bin/env bash
# Round 2, GPU call 11: GEMM + statistics for N = 256 / 384.
set -x
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_linear_stats.py tests/test_gpu_models.py tests/test_gpu_static.py -m gpu -x -q > $O/r2c11_tests.log 2>&1; echo "rc=$?" >> $O/r2c11_tests.log
tail -25 $O/r2c11_tests.log
timeout 600 python bench.py --no-cpu-baseline --no-stock-gpu > $O/r2c11_bench_sswl.json 2> $O/r2c11_bench_sswl.err; grep -E "Error" $O/r2c11_bench_sswl.err | cut -c1-300
timeout 600 python bench.py --batch 128 --no-cpu-baseline --no-stock-gpu --no-roofline > $O/r2c11_bench_sswl128.json 2> $O/r2c11_bench_sswl128.err
python - <<'PY'
import json
for f in ("r2c11_bench_sswl","r2c11_bench_sswl128"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"],3), d["gpu_launches"]/d["steps"], "e2e", round(d["e2e"]["value"]), d["e2e"]["step_ms"])
    except Exception as e: print(f, "ERR", e)
PY
BATCH=1024 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \
  --csv --log-file $O/r2c11_step1024_launches.csv python profiles/run_step.py > $O/r2c11_step1024.log 2>&1
python profiles/launch_summary.py $O/r2c11_step1024_launches.csv 22
