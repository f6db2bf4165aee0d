#!/usr/bin/env bash
# Round 2, GPU call 57: embedding plan in one library call: tests, loader launch list, e2e at 128 / 1024 graphs.
set -x
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2c57_tests.log 2>&1; tail -3 $O/r2c57_tests.log
PYGHO_B200_PROFILE_LOADER=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off \
  --csv --log-file $O/r2c57_loader_b128.csv python bench.py --batch 128 --no-cpu-baseline --no-stock-gpu --no-roofline --steps 2 > $O/r2c57_loader_b128.log 2>&1
python profiles/launch_summary.py $O/r2c57_loader_b128.csv 8
for B in 128 1024; do
  PYGHO_B200_BENCH_TRACE=1 timeout 600 python bench.py --batch $B --no-cpu-baseline --no-stock-gpu --no-roofline > $O/r2c57_bench_b${B}.json 2> $O/r2c57_bench_b${B}.err
  python -c "import json; d=json.loads(open('$O/r2c57_bench_b${B}.json').read().strip().splitlines()[-1]); print('B=$B', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['step_ms'])"
  grep "loader host" $O/r2c57_bench_b${B}.err | cut -c1-160
done
