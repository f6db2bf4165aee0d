#!/usr/bin/env bash
# Round 2, GPU call 22: mamamm algo 4: barrier-wait suspend hint sweep + ablation.
set -x
O=gpurun_out; mkdir -p $O
ABLATE=1 ALGOS=4 ITERS=20 timeout 600 python profiles/run_mamamm.py > $O/r2c22_mamamm_ablate.txt 2>&1; cat $O/r2c22_mamamm_ablate.txt
