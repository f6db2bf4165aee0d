#!/usr/bin/env bash
# Round 2, GPU call 8: TMA-staged mamamm (algo 3), linear_stats test fix.
set -x
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_backend.py -m gpu -x -q -k "mamamm or masked_golden" > $O/r2c8_tests_mamamm.log 2>&1; echo "rc=$?" >> $O/r2c8_tests_mamamm.log
tail -25 $O/r2c8_tests_mamamm.log
ALGOS=0,2,3 ITERS=20 timeout 300 python profiles/run_mamamm.py > $O/r2c8_mamamm_times.txt 2>&1; cat $O/r2c8_mamamm_times.txt
timeout 900 python -m pytest tests -m gpu -q > $O/r2c8_tests.log 2>&1; echo "rc=$?" >> $O/r2c8_tests.log
tail -12 $O/r2c8_tests.log
