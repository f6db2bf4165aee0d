#!/usr/bin/env bash
# Round 2, GPU call 2: ticketed BN / direct gradient accumulation / static-capacity graph feeding.
set -x
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_static.py -m gpu -x -q > $O/r2c2_tests_static.log 2>&1; echo "rc=$?" >> $O/r2c2_tests_static.log
tail -25 $O/r2c2_tests_static.log
python -m pytest tests -m gpu -q --deselect tests/test_gpu_static.py > $O/r2c2_tests.log 2>&1; echo "rc=$?" >> $O/r2c2_tests.log
tail -8 $O/r2c2_tests.log
python bench.py --no-cpu-baseline --no-stock-gpu > $O/r2c2_bench_sswl.json 2> $O/r2c2_bench_sswl.err; tail -c 1500 $O/r2c2_bench_sswl.err
python bench.py --batch 128 --no-cpu-baseline --no-stock-gpu > $O/r2c2_bench_sswl128.json 2> $O/r2c2_bench_sswl128.err; tail -c 600 $O/r2c2_bench_sswl128.err
BATCH=128 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \
  --csv --log-file $O/r2c2_step128_launches.csv python profiles/run_step.py > $O/r2c2_step128.log 2>&1
BATCH=1024 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \
  --csv --log-file $O/r2c2_step1024_launches.csv python profiles/run_step.py > $O/r2c2_step1024.log 2>&1
cat $O/r2c2_bench_sswl.json | head -c 3000; echo; cat $O/r2c2_bench_sswl128.json | head -c 1500
