#!/bin/bash
# round-2 #18: column-wise data entry (clv_set_data_columns) + register-accumulating initialisation kernel:
# parity tests, stage timing of the set-up calls (both entries; CLV_INIT_GENERIC=1 = the previous init kernel), the bench line
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/r02_pytest18.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest18.log
tail -4 $O/r02_pytest18.log
{ echo "== default"; timeout 200 python tools/e2e_stages.py; echo "== CLV_INIT_GENERIC=1"; CLV_INIT_GENERIC=1 timeout 200 python tools/e2e_stages.py; } > $O/r02_e2e_stages18.log 2>&1
cat $O/r02_e2e_stages18.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-ess --no-configs --no-forecast --no-cpu-baseline --no-peaks ) > $O/r02_bench18.json 2> $O/r02_bench18.err; echo "bench rc=$?"
tail -c 400 $O/r02_bench18.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench18.json").read().strip().splitlines()[-1])
print("value %.4g ms/step %.4f launches %d e2e %.4g (%s) pageable %.4g (%s) digest ok %s" % (d["value"], d["ms_per_step"], d["gpu_launches"], d["e2e"]["value"], d["e2e"]["seconds_all_runs"], d["e2e_pageable"]["value"], d["e2e_pageable"]["seconds_all_runs"], d["digest"]["matches_committed"]))
PY
