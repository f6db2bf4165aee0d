#!/usr/bin/env bash
# Round 2, GPU call 28: mamamm algo 4 with one polling lane per warp: tests + scaling.
set -x
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_backend.py -m gpu -x -q -k "mamamm" > $O/r2c28_tests.log 2>&1; tail -3 $O/r2c28_tests.log
timeout 600 python profiles/mamamm_smem_scaling.py > $O/r2c28_mamamm_scaling.txt 2>&1; cat $O/r2c28_mamamm_scaling.txt
