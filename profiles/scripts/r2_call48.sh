#!/usr/bin/env bash
# Round 2, GPU call 48: where to trigger the dependent launch: at the top of every kernel (A) or never
# explicitly = at kernel completion (B, rebuilt on the box with -DPGH_PDL_NO_TRIGGER); same box, back to back.
set -x
O=gpurun_out; mkdir -p $O
meas() {
  python bench.py --batch 128 --no-cpu-baseline --no-stock-gpu --no-roofline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 sswl128', round(d['value']), round(d['ms_per_step'],3))"
  python bench.py --no-cpu-baseline --no-stock-gpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('$1 sswl1024', round(d['value']), round(d['ms_per_step'],3), round(r['frac'],3), round(r['us_per_launch'],2))"
  python bench.py --workload dssgnn_sr25 --no-cpu-baseline --no-stock-gpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('$1 dssgnn', round(d['value']), round(d['ms_per_step'],3), round(r['frac'],3), round(r['us_per_launch'],2))"
}
meas A_top
touch pygho_b200/csrc/common.cuh
EXTRA_FLAGS=-DPGH_PDL_NO_TRIGGER bash pygho_b200/csrc/build.sh | tail -1
meas B_none
PYGHO_B200_NO_PDL=1 meas C_off
