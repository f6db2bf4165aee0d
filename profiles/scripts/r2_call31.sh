#!/usr/bin/env bash
# Round 2, GPU call 31: mamamm algo 4 with 8 x 8 accumulators per lane (1 channel per lane) + LPT queue order.
set -x
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_backend.py -m gpu -x -q -k "mamamm" > $O/r2c31_tests.log 2>&1; tail -5 $O/r2c31_tests.log
timeout 600 python profiles/mamamm_smem_scaling.py > $O/r2c31_mamamm_scaling.txt 2>&1; cat $O/r2c31_mamamm_scaling.txt
