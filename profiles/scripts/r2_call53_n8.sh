#!/usr/bin/env bash
# Round 2, GPU call 53 (8 GPUs): final strong-scaling sweep of the round with the final code (v4 lines) + new test.
set -x
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_backend.py -m gpu -x -q -k "gather_product" 2>&1 | tail -2
run() { # tag, nproc, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $((29530 + $2)) \
    bench.py --gpus $2 --steps 20 --warmup 5 ${@:3} > $O/r2c53_bench_$1.json 2> $O/r2c53_bench_$1.err
  echo "rc=$?"; grep -E "Error|Traceback" $O/r2c53_bench_$1.err | head -5
  python - <<PY
import json
try:
    d=json.loads(open("$O/r2c53_bench_$1.json").read().strip().splitlines()[-1])
    e=d["e2e"]; print("$1", "value", round(d["value"]), "ms", round(d["ms_per_step"],3), "per_gpu", d["config"]["per_gpu_batch"], "| e2e", round(e["value"]), round(e["ms_per_step"],3), e["step_ms"], d["clocks"])
except Exception as ex: print("$1 ERR", ex)
PY
}
run n8 8 --no-cpu-baseline --no-stock-gpu
run n4 4 --no-cpu-baseline --no-stock-gpu
run n2 2 --no-cpu-baseline --no-stock-gpu
timeout 300 python bench.py --no-cpu-baseline --no-stock-gpu > $O/r2c53_bench_n1.json 2> $O/r2c53_bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c53_bench_n1.json").read().strip().splitlines()[-1]); print("n1", round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["clocks"])
PY
