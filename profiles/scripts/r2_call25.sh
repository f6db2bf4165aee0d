#!/usr/bin/env bash
# Round 2, GPU call 25: mamamm algo 4 v5 (two CTAs per SM, pad fill by the producer): tests, graph-replay ablation, trace.
set -x
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_backend.py -m gpu -x -q -k "mamamm" > $O/r2c25_tests.log 2>&1; tail -5 $O/r2c25_tests.log
timeout 600 python profiles/mamamm_smem_trace.py > $O/r2c25_mamamm_trace.txt 2>&1; head -60 $O/r2c25_mamamm_trace.txt
