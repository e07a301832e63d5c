#!/bin/bash
# round-2, 8 GPUs, E32 kernel: sharded == 1-GPU chain (D=2 over peer mailboxes, D=3 over NCCL), the driver's bench command at N=8,
# and a short NCCL-transport line
cd "$(dirname "$0")/.."
G=${1:-8}
O=gpurun_out
mkdir -p $O
L=$O/r02_sharded_check_${G}gpu.log
: > $L
for DP in "2 1" "3 0"; do set -- $DP
  CLV_P2P=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$G --master-addr 127.0.0.1 --master-port 29561 \
     tools/sharded_check.py 2000003 $1 2>&1 | grep -E "SHARDED_OK|Error|error|assert" >> $L
done
cat $L
show() {
  python - <<PY
import json
try:
    d = json.loads(open("$1").read().strip().splitlines()[-1])
    print("$1: N=%d value %.4g ms/step %.4f e2e %.4g stationary %.4g digest %s %s match %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["stationary"]["value"], d["digest"]["level_2_sha256"][:16], d["digest"]["level_1_hash64"], d["digest"].get("matches_committed")))
    if d.get("ess"): print("   ess strong %s weak %s" % (json.dumps(d["ess"].get("strong"))[:300], json.dumps(d["ess"].get("weak"))[:300]))
except Exception as e:
    print("no JSON line:", e)
PY
}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$G --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $G --steps 20 --warmup 5 > $O/r02_bench_${G}gpu.json 2> $O/r02_bench_${G}gpu.err
echo "bench N=$G rc=$?"; tail -c 200 $O/r02_bench_${G}gpu.err; show $O/r02_bench_${G}gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$G --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus $G --steps 200 --warmup 10 --collective nccl \
   --no-ess --no-configs --no-forecast --no-cpu-baseline --no-peaks > $O/r02_bench_${G}gpu_nccl.json 2> $O/r02_bench_${G}gpu_nccl.err
echo "nccl rc=$?"; show $O/r02_bench_${G}gpu_nccl.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$G --master-addr 127.0.0.1 --master-port 29564 bench.py --gpus $G --steps 200 --warmup 10 \
   --no-ess --no-configs --no-forecast --no-cpu-baseline --no-peaks > $O/r02_bench_${G}gpu_200.json 2> $O/r02_bench_${G}gpu_200.err
echo "p2p 200 rc=$?"; show $O/r02_bench_${G}gpu_200.json
