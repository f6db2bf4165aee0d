#!/usr/bin/env bash
# Round 2, GPU call 14: ncu --set full captures of the kernels named in DESIGN.md (one launch each).
set -x
O=gpurun_out; mkdir -p $O
NCU="ncu --set full --clock-control none"
python profiles/run_linear_stats.py > $O/r2c14_linear_stats_times.txt 2>&1; cat $O/r2c14_linear_stats_times.txt
N=384 python profiles/run_linear_stats.py > $O/r2c14_bn384_times.txt 2>&1; cat $O/r2c14_bn384_times.txt
ITERS=4 timeout 300 $NCU -k regex:seg_gmr_lean -s 3 -c 1 -o $O/r2c14_seg_gmr_lean_fwd python profiles/run_spspmm.py > $O/r2c14_ncu1.log 2>&1
ITERS=4 timeout 300 $NCU -k regex:linear_stats -s 3 -c 1 -o $O/r2c14_linear_stats python profiles/run_linear_stats.py > $O/r2c14_ncu2.log 2>&1
ITERS=4 timeout 300 $NCU -k regex:bn_stats_kernel -s 3 -c 1 -o $O/r2c14_bn_stats python profiles/run_linear_stats.py > $O/r2c14_ncu3.log 2>&1
ITERS=4 N=384 timeout 300 $NCU -k regex:bn_act_bwd_apply -s 3 -c 1 -o $O/r2c14_bn_apply384 python profiles/run_linear_stats.py > $O/r2c14_ncu4.log 2>&1
ALGOS=2 ITERS=4 timeout 300 $NCU -k regex:mamamm_tc_pipe -s 10 -c 1 -o $O/r2c14_mamamm_pipe_ext python profiles/run_mamamm.py > $O/r2c14_ncu5.log 2>&1
ITERS=4 KEY=X___X___1___X___0 timeout 300 $NCU -k regex:seg_gmr_lean -s 3 -c 1 -o $O/r2c14_seg_gmr_2fwl python profiles/run_spspmm.py > $O/r2c14_ncu6.log 2>&1
for r in seg_gmr_lean_fwd linear_stats bn_stats bn_apply384 mamamm_pipe_ext seg_gmr_2fwl; do python profiles/ncu_summary.py $O/r2c14_$r.ncu-rep > $O/r2c14_$r.summary.txt 2>&1; head -24 $O/r2c14_$r.summary.txt; done
ls -la $O/*.ncu-rep | tail -8
