#!/usr/bin/env bash
# Round 2, GPU call 37: warm-cache launch list of a 128-graph step (ncu --cache-control none): where the 2.65 ms go.
set -x
O=gpurun_out; mkdir -p $O
BATCH=128 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off \
  --csv --log-file $O/r2c37_step128_warm_launches.csv python profiles/run_step.py > $O/r2c37_step128.log 2>&1
python profiles/launch_summary.py $O/r2c37_step128_warm_launches.csv 30
