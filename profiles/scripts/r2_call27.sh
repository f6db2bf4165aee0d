#!/usr/bin/env bash
# Round 2, GPU call 27: mamamm algo 4 time against batch size (fixed vs per-unit cost).
set -x
O=gpurun_out; mkdir -p $O
timeout 600 python profiles/mamamm_smem_scaling.py > $O/r2c27_mamamm_scaling.txt 2>&1; cat $O/r2c27_mamamm_scaling.txt
