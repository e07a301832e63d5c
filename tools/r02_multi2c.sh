#!/bin/bash
# round-2, 2 GPUs, after the column-wise data entry / new initialisation kernel: multi-GPU tests, sharded == 1-GPU chain,
# the driver's bench command at N=2, and the 8-GPU shard size on 2 GPUs over peer mailboxes
cd "$(dirname "$0")/.."
G=${1:-2}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/r02_pytest_multi_${G}gpu_c.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_multi_${G}gpu_c.log
tail -3 $O/r02_pytest_multi_${G}gpu_c.log
L=$O/r02_sharded_check_${G}gpu_c.log
: > $L
for DP in "2 1" "3 0"; do set -- $DP
  CLV_P2P=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$G --master-addr 127.0.0.1 --master-port 29551 \
     tools/sharded_check.py 1000003 $1 2>&1 | grep -E "SHARDED_OK|Error|error|assert" >> $L
done
cat $L
show() {
  python - <<PY
import json
try:
    d = json.loads(open("$1").read().strip().splitlines()[-1])
    print("$1: N=%d value %.4g ms/step %.4f e2e %.4g digest %s %s match %s launches %d" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["digest"]["level_2_sha256"][:16], d["digest"]["level_1_hash64"], d["digest"].get("matches_committed"), d["gpu_launches"]))
except Exception as e:
    print("no JSON line:", e)
PY
}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$G --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $G --steps 20 --warmup 5 > $O/r02_bench_${G}gpu_c.json 2> $O/r02_bench_${G}gpu_c.err
echo "bench N=$G rc=$?"; tail -c 300 $O/r02_bench_${G}gpu_c.err; show $O/r02_bench_${G}gpu_c.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$G --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus $G --steps 200 --warmup 10 \
   --customers $((1250000 * G)) --collective p2p --no-ess --no-configs --no-forecast --no-cpu-baseline --no-peaks > $O/r02_bench_shard8_${G}gpu_p2p_c.json 2> $O/r02_bench_shard8_${G}gpu_p2p_c.err
echo "shard8 p2p rc=$?"; show $O/r02_bench_shard8_${G}gpu_p2p_c.json
