#!/usr/bin/env bash
# Round 2, GPU call 19: ablation of the mamamm algo-4 kernel (which phase bounds it).
set -x
O=gpurun_out; mkdir -p $O
ABLATE=1 ALGOS=4 ITERS=20 timeout 600 python profiles/run_mamamm.py > $O/r2c19_mamamm_ablate.txt 2>&1; cat $O/r2c19_mamamm_ablate.txt
