#!/usr/bin/env bash
# Round 2, GPU call 50: compute-sanitizer memcheck over the tests of the kernels added / changed late in the round.
set -x
O=gpurun_out; mkdir -p $O
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_backend.py tests/test_gpu_static.py -m gpu -x -q \
  -k "masked_pool or mamamm_smem or acd_regroup or sswl_merged or fused_epilogue or static or feeder" > $O/r2c50_memcheck.log 2>&1
echo "rc=$?" >> $O/r2c50_memcheck.log
tail -12 $O/r2c50_memcheck.log
