#!/usr/bin/env bash
# Round 2, GPU call 12 (8 GPUs): strong scaling at the BASELINE config (global batch 1024).
set -x
O=gpurun_out; mkdir -p $O
nvidia-smi -L | head -8
export PYGHO_B200_BENCH_TRACE=1
run() { # tag, nproc, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $((29520 + $2)) \
    bench.py --gpus $2 --steps 20 --warmup 5 ${@:3} > $O/r2c12_bench_$1.json 2> $O/r2c12_bench_$1.err
  echo "rc=$?"; grep -E "Error|error|Traceback" $O/r2c12_bench_$1.err | head -5
  python - <<PY
import json
try:
    d=json.loads(open("$O/r2c12_bench_$1.json").read().strip().splitlines()[-1])
    e=d["e2e"]; print("$1", "value", round(d["value"]), "ms", round(d["ms_per_step"],3), "per_gpu", d["config"]["per_gpu_batch"], "| e2e", round(e["value"]), round(e["ms_per_step"],3), e["step_ms"])
except Exception as ex: print("$1 ERR", ex)
PY
}
run n8 8
run n8_syncbn 8 --syncbn
run n4 4
run n8_weak 8 --scaling weak --no-roofline
grep "trace" $O/r2c12_bench_n8.err | head -4 | cut -c1-300
