#!/bin/bash
# round-2 A/B #22: non-temporal stores in the host copy pool (CLV_COPY_NT=0 = plain memcpy) x ring slot size
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
{
for cfg in "0 16" "1 16" "0 64" "1 64" "1 16" "0 16"; do set -- $cfg
  echo "== CLV_COPY_NT=$1 CLV_STAGE_PIECE_MB=$2"
  CLV_COPY_NT=$1 CLV_STAGE_PIECE_MB=$2 timeout 300 python tools/c2_timing.py 2>&1 | grep "store_level1=True"
done
for nt in 0 1; do
  echo "== configs CLV_COPY_NT=$nt"
  CLV_COPY_NT=$nt timeout 300 python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import bench
out = bench.configs_block(0, with_cpu=False)
print({k: ([round(w, 3) for w in v["wall_s_all_runs"]], round(v["level_1_to_host_GB"], 2)) for k, v in out.items()})
PY
done
echo "== e2e stages (H2D through the ring), CLV_COPY_NT=0 then 1"
CLV_COPY_NT=0 timeout 200 python tools/e2e_stages.py | grep pageable
CLV_COPY_NT=1 timeout 200 python tools/e2e_stages.py | grep pageable
} > $O/r02_ab22.log 2>&1
cat $O/r02_ab22.log
