#!/bin/bash
# round-2, 4 GPUs, final state: sharded == 1-GPU chain over peer mailboxes, and the driver's bench command at N=4
cd "$(dirname "$0")/.."
G=4
O=gpurun_out
mkdir -p $O
L=$O/r02_sharded_check_${G}gpu.log
CLV_P2P=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$G --master-addr 127.0.0.1 --master-port 29571 \
   tools/sharded_check.py 1000003 2 2>&1 | grep -E "SHARDED_OK|Error|error|assert" > $L
cat $L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$G --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus $G --steps 20 --warmup 5 > $O/r02_bench_${G}gpu_c.json 2> $O/r02_bench_${G}gpu_c.err
echo "bench N=$G rc=$?"; tail -c 200 $O/r02_bench_${G}gpu_c.err
python - <<PY
import json
d = json.loads(open("$O/r02_bench_${G}gpu_c.json").read().strip().splitlines()[-1])
print("N=%d value %.4g ms/step %.4f e2e %.4g stationary %.4g digest %s %s match %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["stationary"]["value"], d["digest"]["level_2_sha256"][:16], d["digest"]["level_1_hash64"], d["digest"].get("matches_committed")))
if d.get("ess"): print("   ess strong %s" % json.dumps(d["ess"].get("strong"))[:300])
PY
