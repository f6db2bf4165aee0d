#!/usr/bin/env bash
# Round 2, GPU call 52: fused tuple initialisation (ops.GatherProduct): full GPU tests, 1024- and 128-graph steps.
set -x
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2c52_tests.log 2>&1; tail -4 $O/r2c52_tests.log
for B in 1024 128; do
  timeout 600 python bench.py --batch $B --no-cpu-baseline --no-stock-gpu --no-roofline > $O/r2c52_bench_b${B}.json 2> $O/r2c52_bench_b${B}.err
  python -c "import json; d=json.loads(open('$O/r2c52_bench_b${B}.json').read().strip().splitlines()[-1]); print('B=$B', d['value'], d['ms_per_step'], d['gpu_launches'], d['e2e']['value'], d['e2e']['step_ms'])"
  grep -E "Error|Traceback" $O/r2c52_bench_b${B}.err | head -3
done
