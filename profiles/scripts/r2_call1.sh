#!/usr/bin/env bash
# Round 2, GPU call 1: parity suite, bench lines of every workload, launch list of a 128-graph
# step (the per-GPU work of the N=8 strong-scaling point), compute-sanitizer over the kernels.
set -x
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2c1_gpu.txt
python -m pytest tests -m gpu -x -q > $O/r2c1_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2c1_tests.log
tail -3 $O/r2c1_tests.log
python bench.py > $O/r2c1_bench_sswl.json 2> $O/r2c1_bench_sswl.err; tail -c 600 $O/r2c1_bench_sswl.err
python bench.py --batch 128 --no-cpu-baseline > $O/r2c1_bench_sswl128.json 2> $O/r2c1_bench_sswl128.err
for wl in ppgn_dd dssgnn_sr25 i2_sr25; do
  timeout 600 python bench.py --workload $wl > $O/r2c1_bench_$wl.json 2> $O/r2c1_bench_$wl.err; tail -c 400 $O/r2c1_bench_$wl.err
done
BATCH=128 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \
  --csv --log-file $O/r2c1_step128_launches.csv python profiles/run_step.py > $O/r2c1_step128.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_backend.py -m gpu -x -q \
  -k "seg_gmr_forward_backward_vs_torch and (128 or 384 or 12) or mamamm or masked_pool or fused" \
  > $O/r2c1_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/r2c1_sanitizer_memcheck.log
tail -5 $O/r2c1_sanitizer_memcheck.log
cat $O/r2c1_bench_sswl.json | head -c 2500
