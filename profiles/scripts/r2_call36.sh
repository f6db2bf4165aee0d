#!/usr/bin/env bash
# Round 2, GPU call 36: programmatic dependent launch for the step kernels: tests, 128- and 1024-graph steps with and without.
set -x
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2c36_tests.log 2>&1; tail -3 $O/r2c36_tests.log
for B in 128 1024; do
  timeout 600 python bench.py --batch $B --no-cpu-baseline --no-stock-gpu --no-roofline > $O/r2c36_bench_b${B}_pdl.json 2> $O/r2c36_bench_b${B}_pdl.err
  python -c "import json; d=json.loads(open('$O/r2c36_bench_b${B}_pdl.json').read().strip().splitlines()[-1]); print('PDL  B=$B', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['step_ms'])"
  PYGHO_B200_NO_PDL=1 timeout 600 python bench.py --batch $B --no-cpu-baseline --no-stock-gpu --no-roofline > $O/r2c36_bench_b${B}_nopdl.json 2> $O/r2c36_bench_b${B}_nopdl.err
  python -c "import json; d=json.loads(open('$O/r2c36_bench_b${B}_nopdl.json').read().strip().splitlines()[-1]); print('noPDL B=$B', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['step_ms'])"
done
