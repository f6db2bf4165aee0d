#!/usr/bin/env bash
# Round 2, GPU call 55: launch list of ONE batch load of the static feeder (128 and 1024 graphs).
set -x
O=gpurun_out; mkdir -p $O
for B in 128 1024; do
PYGHO_B200_PROFILE_LOADER=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off \
  --csv --log-file $O/r2c55_loader_b${B}.csv python bench.py --batch $B --no-cpu-baseline --no-stock-gpu --no-roofline --steps 2 > $O/r2c55_loader_b${B}.log 2>&1
python profiles/launch_summary.py $O/r2c55_loader_b${B}.csv 16
done
