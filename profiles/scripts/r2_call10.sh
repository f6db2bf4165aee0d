#!/usr/bin/env bash
# Round 2, GPU call 10: full parity suite after the clean-up, every workload's bench line.
set -x
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/r2c10_tests.log 2>&1; echo "rc=$?" >> $O/r2c10_tests.log
tail -15 $O/r2c10_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2c10_smoke.log 2>&1; tail -2 $O/r2c10_smoke.log
timeout 900 python bench.py > $O/r2c10_bench_sswl.json 2> $O/r2c10_bench_sswl.err
for wl in ppgn_dd dssgnn_sr25 i2_sr25; do
  timeout 900 python bench.py --workload $wl > $O/r2c10_bench_$wl.json 2> $O/r2c10_bench_$wl.err; tail -c 300 $O/r2c10_bench_$wl.err
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2c10_bench_ref.json 2> $O/r2c10_bench_ref.err
python - <<'PY'
import json
for f in ("sswl","ppgn_dd","dssgnn_sr25","i2_sr25","ref"):
    try:
        d=json.loads(open(f"gpurun_out/r2c10_bench_{f}.json").read().strip().splitlines()[-1])
        e=d.get("e2e") or {}
        print(f, round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(e.get("value",0),1), e.get("mode","")[:40], "roof", (d.get("roofline") or {}).get("frac"), "stock", (d.get("stock_gpu_baseline") or {}).get("own_over_stock"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as ex: print(f, "ERR", ex)
PY
