#!/bin/bash
# round-2 A/B #24: first touch of the caller's level-1 array by host threads while the sweeps run (CLV_FIRST_TOUCH=0 = before)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
{
for ft in 0 1; do
  echo "== CLV_FIRST_TOUCH=$ft"
  CLV_FIRST_TOUCH=$ft timeout 300 python tools/d2h_probe.py 2>&1 | grep -v madvise
  CLV_FIRST_TOUCH=$ft timeout 300 python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import bench
out = bench.configs_block(0, with_cpu=False)
print({k: ([round(w, 3) for w in v["wall_s_all_runs"]], round(v["level_1_to_host_GB"], 2)) for k, v in out.items()})
PY
done
for nt in 4 16; do
  echo "== CLV_FIRST_TOUCH_THREADS=$nt"
  CLV_FIRST_TOUCH_THREADS=$nt timeout 300 python tools/d2h_probe.py 2>&1 | grep "fresh np.empty"
done
} > $O/r02_ab24.log 2>&1
cat $O/r02_ab24.log
for ft in 0 1; do
CLV_FIRST_TOUCH=$ft timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-ess --no-configs --no-forecast --no-cpu-baseline --no-peaks > $O/r02_bench24_$ft.json 2> $O/r02_bench24_$ft.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_bench24_$ft.json").read().strip().splitlines()[-1])
print("first touch $ft: value %.4g e2e %.4g (%s) pageable %.4g (%s) digest ok %s" % (d["value"], d["e2e"]["value"], d["e2e"]["seconds_all_runs"], d["e2e_pageable"]["value"], d["e2e_pageable"]["seconds_all_runs"], d["digest"]["matches_committed"]))
PY
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "staged or caller_provided or persistent or api_layout" > $O/r02_pytest24.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest24.log; tail -3 $O/r02_pytest24.log
