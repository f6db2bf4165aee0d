#!/usr/bin/env bash
# Round 2, GPU call 15: staged segmented-reduce kernel.
set -x
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_backend.py -m gpu -x -q -k "staged or seg_gmr" > $O/r2c15_tests_staged.log 2>&1; echo "rc=$?" >> $O/r2c15_tests_staged.log
tail -15 $O/r2c15_tests_staged.log
timeout 600 python profiles/run_staged.py > $O/r2c15_staged_times.txt 2>&1; cat $O/r2c15_staged_times.txt
timeout 900 python -m pytest tests -m gpu -q > $O/r2c15_tests.log 2>&1; echo "rc=$?" >> $O/r2c15_tests.log; tail -6 $O/r2c15_tests.log
for wl in dssgnn_sr25 i2_sr25; do
  timeout 900 python bench.py --workload $wl --no-cpu-baseline --no-stock-gpu > $O/r2c15_bench_$wl.json 2> $O/r2c15_bench_$wl.err; tail -c 300 $O/r2c15_bench_$wl.err
done
python - <<'PY'
import json
for f in ("dssgnn_sr25","i2_sr25"):
    try:
        d=json.loads(open(f"gpurun_out/r2c15_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "roof", d["roofline"].get("frac"), "spspmm", (d.get("roofline_spspmm") or {}).get("us_per_launch"), (d.get("roofline_spspmm") or {}).get("frac"))
    except Exception as ex: print(f, "ERR", ex)
PY
