#!/usr/bin/env bash
# Round 2, GPU call 39: op-level roofline table v2 + ncu --set full of the default mamamm kernel (algo 4, largest first)
# and of the warp-cooperative masked pooling kernel.
set -x
O=gpurun_out; mkdir -p $O
timeout 900 python profiles/run_ops.py --md $O/r2c39_op_rooflines.md > $O/r2c39_run_ops.log 2>&1; tail -3 $O/r2c39_run_ops.log
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:mamamm_smem -s 60 -c 1 -o $O/r2c39_mamamm_smem_lpt python profiles/run_masked.py --mamamm > $O/r2c39_ncu1.log 2>&1
timeout 300 $NCU -k regex:masked_pool_warp -s 5 -c 1 -o $O/r2c39_masked_pool_warp python profiles/run_masked.py > $O/r2c39_ncu2.log 2>&1
for r in mamamm_smem_lpt masked_pool_warp; do python profiles/ncu_summary.py $O/r2c39_$r.ncu-rep > $O/r2c39_$r.summary.txt 2>&1; head -30 $O/r2c39_$r.summary.txt; done
