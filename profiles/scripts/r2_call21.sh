#!/usr/bin/env bash
# Round 2, GPU call 21: mamamm algo 4 v3 (pad fill on the bulk-copy engine): tests, ablation, ncu capture.
set -x
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_backend.py -m gpu -x -q -k "mamamm" > $O/r2c21_tests.log 2>&1; tail -5 $O/r2c21_tests.log
ABLATE=1 ALGOS=2,4 ITERS=20 timeout 600 python profiles/run_mamamm.py > $O/r2c21_mamamm_ablate.txt 2>&1; cat $O/r2c21_mamamm_ablate.txt
ALGOS=4 ITERS=2 timeout 300 ncu --set full --import-source on --clock-control none -k regex:mamamm_smem -s 6 -c 1 -o $O/r2c21_mamamm_smem_ext python profiles/run_mamamm.py > $O/r2c21_ncu.log 2>&1
python profiles/ncu_summary.py $O/r2c21_mamamm_smem_ext.ncu-rep > $O/r2c21_mamamm_smem_ext.summary.txt 2>&1; cat $O/r2c21_mamamm_smem_ext.summary.txt
python profiles/ncu_source.py $O/r2c21_mamamm_smem_ext.ncu-rep 40 > $O/r2c21_mamamm_smem_ext.source.txt 2>&1; head -80 $O/r2c21_mamamm_smem_ext.source.txt
