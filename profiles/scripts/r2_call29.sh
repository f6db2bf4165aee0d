#!/usr/bin/env bash
# Round 2, GPU call 29: warp-cooperative masked pooling (ballot mask bits): tests + timings.
set -x
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_backend.py tests/test_gpu_models.py -m gpu -x -q -k "pool or masked or ppgn or PPGN or dense" > $O/r2c29_tests.log 2>&1; tail -3 $O/r2c29_tests.log
timeout 600 python profiles/run_masked.py --mamamm > $O/r2c29_masked.txt 2>&1; cat $O/r2c29_masked.txt
