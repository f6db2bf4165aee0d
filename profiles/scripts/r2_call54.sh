#!/usr/bin/env bash
# Round 2, GPU call 54: what the driver runs at round end, on the final code: pytest -m gpu, smoke(), bench.py (default flags).
set -x
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests/ -x -q -m gpu > $O/r2c54_tests.log 2>&1; tail -3 $O/r2c54_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
/usr/bin/time -v timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2c54_bench.json 2> $O/r2c54_bench.err; grep -E "Elapsed|Error|Traceback" $O/r2c54_bench.err | head -5
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c54_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("metric","value","ms_per_step","scaling","gpu_launches","clocks")})
print("e2e", d["e2e"]["value"], d["e2e"]["step_ms"]); print("roofline", d["roofline"]["frac"], d["roofline"]["us_per_launch"]); print("cpu", d["cpu_baseline"]); print("stock", d["stock_gpu_baseline"]["value"], d["stock_gpu_baseline"]["own_over_stock"])
PY
