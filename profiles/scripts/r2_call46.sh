#!/usr/bin/env bash
# Round 2, GPU call 46: batched device copies (pgh_multi_copy) in the static feeder: tests, e2e at 1024 and 128 graphs.
set -x
O=gpurun_out; mkdir -p $O
for B in 1024 128; do
  PYGHO_B200_BENCH_TRACE=1 timeout 600 python bench.py --batch $B --no-cpu-baseline --no-stock-gpu --no-roofline > $O/r2c46_bench_b${B}.json 2> $O/r2c46_bench_b${B}.err
  python -c "import json; d=json.loads(open('$O/r2c46_bench_b${B}.json').read().strip().splitlines()[-1]); print('B=$B', d['value'], d['ms_per_step'], d['gpu_launches'], d['e2e']['value'], d['e2e']['step_ms'])"
  grep -E "Error|Traceback" $O/r2c46_bench_b${B}.err | head -3
done
grep "loader host" $O/r2c46_bench_b*.err | cut -c1-200
