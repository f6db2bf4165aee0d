#!/usr/bin/env bash
# Round 2, GPU call 32: timeline of the 8 x 8-tile mamamm algo 4 (CTA 0 producer + consumer warp 0).
set -x
O=gpurun_out; mkdir -p $O
TRACE_ONLY=1 timeout 300 python profiles/mamamm_smem_trace.py > $O/r2c32_trace.txt 2>&1
TRACE_ONLY=1 TRACE_DBG=4 timeout 300 python profiles/mamamm_smem_trace.py > $O/r2c32_trace_nostore.txt 2>&1
head -5 $O/r2c32_trace.txt
