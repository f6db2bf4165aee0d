#!/usr/bin/env bash
# Round 2, GPU call 18: mamamm algo 4 (TMA-fed smem ring, fp32) correctness + timing; FlatAdamW; bench line.
set -x
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_backend.py tests/test_gpu_static.py tests/test_gpu_linear_stats.py -m gpu -x -q -k "mamamm or adamw or mlp_block" > $O/r2c18_tests.log 2>&1; tail -8 $O/r2c18_tests.log
ALGOS=0,2,4 ITERS=20 timeout 600 python profiles/run_mamamm.py > $O/r2c18_mamamm.txt 2>&1; cat $O/r2c18_mamamm.txt
timeout 900 python bench.py --steps 20 --warmup 5 --no-stock-gpu > $O/r2c18_bench_n1.json 2> $O/r2c18_bench_n1.err; tail -c 1500 $O/r2c18_bench_n1.json
PYGHO_B200_MAMAMM_ALGO=4 timeout 900 python bench.py --workload ppgn_dd --steps 20 --warmup 5 --no-stock-gpu > $O/r2c18_bench_ppgn_algo4.json 2> $O/r2c18_bench_ppgn_algo4.err; tail -c 1200 $O/r2c18_bench_ppgn_algo4.json
