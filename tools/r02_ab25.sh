#!/bin/bash
# round-2 A/B #25: first-touch threads: how many (C1/C2/C3 through the drop-in modules, two runs each), and on the forecast's
# output arrays (clv_forecast, 1 M customers x 64 draws, pageable host arrays in and out)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
{
for nt in 8 12 16 6; do
  echo "== CLV_FIRST_TOUCH_THREADS=$nt"
  CLV_FIRST_TOUCH_THREADS=$nt timeout 300 python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import bench
out = bench.configs_block(0, with_cpu=False)
print({k: ([round(w, 3) for w in v["wall_s_all_runs"]], round(v["level_1_to_host_GB"], 2)) for k, v in out.items()})
PY
done
for ft in 0 1; do
  echo "== forecast API path, CLV_FIRST_TOUCH=$ft"
  CLV_FIRST_TOUCH=$ft timeout 300 python tools/forecast_api_timing.py
done
} > $O/r02_ab25.log 2>&1
cat $O/r02_ab25.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "forecast or staged" > $O/r02_pytest25.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest25.log; tail -3 $O/r02_pytest25.log
