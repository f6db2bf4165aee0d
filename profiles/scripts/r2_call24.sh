#!/usr/bin/env bash
# Round 2, GPU call 24: mamamm algo 4 timeline trace + graph-replay timing.
set -x
O=gpurun_out; mkdir -p $O
timeout 600 python profiles/mamamm_smem_trace.py > $O/r2c24_mamamm_trace.txt 2>&1; head -150 $O/r2c24_mamamm_trace.txt
