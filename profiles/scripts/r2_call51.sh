#!/usr/bin/env bash
# Round 2, GPU call 51: block-tiled segmented reduce: test + timings against the streaming kernels.
set -x
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_backend.py -m gpu -x -q -k "block_tiled or seg_gmr or spspmm or full_size" > $O/r2c51_tests.log 2>&1; tail -5 $O/r2c51_tests.log
timeout 600 python profiles/run_block_tiles.py > $O/r2c51_block_tiles.txt 2>&1; cat $O/r2c51_block_tiles.txt
