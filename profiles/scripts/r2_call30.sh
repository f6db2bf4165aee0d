#!/usr/bin/env bash
# Round 2, GPU call 30: split-row masked pooling (one CTA per output row when few rows, long extent).
set -x
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_backend.py -m gpu -x -q -k "pool or masked" > $O/r2c30_tests.log 2>&1; tail -3 $O/r2c30_tests.log
timeout 600 python profiles/run_masked.py > $O/r2c30_masked.txt 2>&1; cat $O/r2c30_masked.txt
