#!/bin/bash
# round-2 A/B #20: MADV_HUGEPAGE on the caller's fresh output arrays (CLV_HUGEPAGE=0 = before): the reference's full-data
# run through Sampler.run (12 GB of level-1 draws), C1/C2/C3 through the drop-in modules, the API-layout forecast
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
{
cat /sys/kernel/mm/transparent_hugepage/enabled
for hp in 0 1; do
  echo "== CLV_HUGEPAGE=$hp"
  CLV_HUGEPAGE=$hp timeout 300 python tools/c2_timing.py
  CLV_HUGEPAGE=$hp timeout 300 python - <<'PY'
import json, sys, os
sys.path.insert(0, os.getcwd())
import bench
out = bench.configs_block(0, with_cpu=False)
print({k: ([round(w, 3) for w in v["wall_s_all_runs"]], round(v["level_1_to_host_GB"], 2)) for k, v in out.items()})
PY
done
} > $O/r02_ab20.log 2>&1
cat $O/r02_ab20.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "staged or forecast or api_layout or persistent" > $O/r02_pytest20.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest20.log; tail -3 $O/r02_pytest20.log
