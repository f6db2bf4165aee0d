#!/usr/bin/env bash
# Round 2, GPU call 35: algo 4 default with LPT order: full GPU test suite, ppgn_dd bench, masked op table, LPT trace.
set -x
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2c35_tests.log 2>&1; tail -5 $O/r2c35_tests.log
timeout 600 python bench.py --workload ppgn_dd > $O/r2c35_bench_ppgn.json 2> $O/r2c35_bench_ppgn.err; cat $O/r2c35_bench_ppgn.json; tail -3 $O/r2c35_bench_ppgn.err
TRACE_ONLY=1 TRACE_LPT=1 timeout 300 python profiles/mamamm_smem_trace.py > $O/r2c35_trace_lpt.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
