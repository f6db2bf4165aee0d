#!/usr/bin/env bash
# Round 2, GPU call 5 (2 GPUs): strong scaling N=2 (global batch 1024 -> 512/GPU), SyncBN variant,
# 2-rank NCCL SyncBN parity test.
set -x
O=gpurun_out; mkdir -p $O
nvidia-smi -L
python -m pytest tests/test_gpu_static.py -m gpu -x -q -k syncbn > $O/r2c5_syncbn_nccl.log 2>&1; echo "rc=$?" >> $O/r2c5_syncbn_nccl.log; tail -4 $O/r2c5_syncbn_nccl.log
export PYGHO_B200_BENCH_TRACE=1
for extra in "" "--syncbn"; do
  tag=n2${extra#--}
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 5 $extra > $O/r2c5_bench_$tag.json 2> $O/r2c5_bench_$tag.err
  echo "rc=$?"; grep -E "trace|Error|error" $O/r2c5_bench_$tag.err | head; cat $O/r2c5_bench_$tag.json | head -c 1800; echo
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 20 --warmup 5 --batch 256 --no-roofline > $O/r2c5_bench_n2_b256.json 2> $O/r2c5_bench_n2_b256.err
cat $O/r2c5_bench_n2_b256.json | head -c 1500
