#!/usr/bin/env bash
# Round 2, GPU call 42: launch list of one 1024-graph step after the merged gradient plan (cold-cache ncu, as before).
set -x
O=gpurun_out; mkdir -p $O
BATCH=1024 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \
  --csv --log-file $O/r2c42_step1024_launches.csv python profiles/run_step.py > $O/r2c42_step1024.log 2>&1
python profiles/launch_summary.py $O/r2c42_step1024_launches.csv 24
