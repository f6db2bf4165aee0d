#!/usr/bin/env bash
# Round 2, GPU call 26: mamamm algo 4 timelines (full kernel and the queue + barriers skeleton).
set -x
O=gpurun_out; mkdir -p $O
TRACE_ONLY=1 TRACE_DBG=15 timeout 600 python profiles/mamamm_smem_trace.py > $O/r2c26_trace_skeleton.txt 2>&1; head -120 $O/r2c26_trace_skeleton.txt
timeout 600 python profiles/mamamm_smem_trace.py > $O/r2c26_trace_full.txt 2>&1; head -14 $O/r2c26_trace_full.txt
