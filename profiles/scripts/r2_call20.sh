#!/usr/bin/env bash
# Round 2, GPU call 20: mamamm algo 4 v2 (producer look-ahead, pad-fill warp, mask prefetch): tests + ablation.
set -x
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_backend.py -m gpu -x -q -k "mamamm" > $O/r2c20_tests.log 2>&1; tail -5 $O/r2c20_tests.log
ABLATE=1 ALGOS=2,4 ITERS=20 timeout 600 python profiles/run_mamamm.py > $O/r2c20_mamamm_ablate.txt 2>&1; cat $O/r2c20_mamamm_ablate.txt
