#!/usr/bin/env bash
# Round 2, GPU call 40: ticket fast paths (one CTA / one group): BN tests, small-size kernel times, 128-graph step.
set -x
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2c40_tests.log 2>&1; tail -3 $O/r2c40_tests.log
python profiles/bn_small.py 2>&1 | grep "ctas/sm= 4" | tee $O/r2c40_bn_small.txt
timeout 600 python bench.py --batch 128 --no-cpu-baseline --no-stock-gpu --no-roofline > $O/r2c40_bench_b128.json 2> $O/r2c40_bench_b128.err
python -c "import json; d=json.loads(open('$O/r2c40_bench_b128.json').read().strip().splitlines()[-1]); print('B=128', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['step_ms'])"
