#!/usr/bin/env bash
# Round 2, GPU call 17: op-level roofline table of the shipped kernels; memcheck over the new kernels.
set -x
O=gpurun_out; mkdir -p $O
timeout 900 python profiles/run_ops.py --md $O/r2c17_op_rooflines.md > $O/r2c17_run_ops.log 2>&1; tail -70 $O/r2c17_run_ops.log | cut -c1-200
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_linear_stats.py tests/test_gpu_static.py -m gpu -x -q \
  -k "not syncbn_sharded and not 300000 and not 70001" > $O/r2c17_sanitizer_memcheck_new.log 2>&1; echo "memcheck rc=$?" >> $O/r2c17_sanitizer_memcheck_new.log
tail -6 $O/r2c17_sanitizer_memcheck_new.log
