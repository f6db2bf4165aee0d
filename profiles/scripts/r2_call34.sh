#!/usr/bin/env bash
# Round 2, GPU call 34: 4 x 4 x 4-channel tiles again, unit = (graph, slab) with inner row passes, LPT order, register-count trace.
set -x
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_backend.py -m gpu -x -q -k "mamamm" > $O/r2c34_tests.log 2>&1; tail -5 $O/r2c34_tests.log
timeout 600 python profiles/mamamm_smem_scaling.py > $O/r2c34_mamamm_scaling.txt 2>&1; cat $O/r2c34_mamamm_scaling.txt
TRACE_ONLY=1 timeout 300 python profiles/mamamm_smem_trace.py > $O/r2c34_trace.txt 2>&1
