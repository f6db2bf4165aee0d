#!/usr/bin/env bash
# Round 2, GPU call 6: first run of the TMA/tcgen05 GEMM with statistics epilogue.
set -x
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_linear_stats.py -m gpu -x -q > $O/r2c6_tests_ls.log 2>&1; echo "rc=$?" >> $O/r2c6_tests_ls.log
tail -40 $O/r2c6_tests_ls.log
timeout 300 python -m pytest tests/test_gpu_static.py -m gpu -x -q > $O/r2c6_tests_static.log 2>&1; echo "rc=$?" >> $O/r2c6_tests_static.log; tail -5 $O/r2c6_tests_static.log
export PYGHO_B200_BENCH_TRACE=1
timeout 600 python bench.py --no-cpu-baseline --no-stock-gpu > $O/r2c6_bench_sswl.json 2> $O/r2c6_bench_sswl.err; grep -E "trace|Error" $O/r2c6_bench_sswl.err | cut -c1-300
PYGHO_B200_FUSED_GEMM=0 timeout 600 python bench.py --no-cpu-baseline --no-stock-gpu --no-roofline > $O/r2c6_bench_sswl_nofused.json 2> $O/r2c6_bench_sswl_nofused.err
python - <<'PY'
import json
for f in ("r2c6_bench_sswl","r2c6_bench_sswl_nofused"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"],3), d["gpu_launches"]/d["steps"], "e2e", round(d["e2e"]["value"]), d["e2e"]["step_ms"])
    except Exception as e: print(f, "ERR", e)
PY
BATCH=1024 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \
  --csv --log-file $O/r2c6_step1024_launches.csv python profiles/run_step.py > $O/r2c6_step1024.log 2>&1
python profiles/launch_summary.py $O/r2c6_step1024_launches.csv 14
