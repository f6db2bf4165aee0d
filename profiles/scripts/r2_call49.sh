#!/usr/bin/env bash
# Round 2, GPU call 49: launch list of one dense PPGN step (128 graphs, n <= 36) + LPT mamamm ncu capture.
set -x
O=gpurun_out; mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \
  --csv --log-file $O/r2c49_step_ppgn_launches.csv python profiles/run_step_dense.py > $O/r2c49_step_ppgn.log 2>&1
python profiles/launch_summary.py $O/r2c49_step_ppgn_launches.csv 30
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mamamm_smem -s 80 -c 1 -o $O/r2c49_mamamm_smem_lpt python profiles/run_masked.py --mamamm > $O/r2c49_ncu1.log 2>&1
python profiles/ncu_summary.py $O/r2c49_mamamm_smem_lpt.ncu-rep > $O/r2c49_mamamm_smem_lpt.summary.txt 2>&1; head -16 $O/r2c49_mamamm_smem_lpt.summary.txt
