#!/bin/bash
# round-2, 2 GPUs, final kernel (two customers per thread + dynamic tiles): sharded == 1-GPU chain, multi-GPU tests,
# bench at N=2, and the 8-GPU shard size (1.25 M customers per GPU) on 2 GPUs
cd "$(dirname "$0")/.."
G=${1:-2}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/r02_pytest_multi_${G}gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_multi_${G}gpu.log
tail -3 $O/r02_pytest_multi_${G}gpu.log
L=$O/r02_sharded_check_${G}gpu.log
: > $L
for D in 2 3; do for P in 1 0; do
  CLV_P2P=$P timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$G --master-addr 127.0.0.1 --master-port 29551 \
     tools/sharded_check.py 1000003 $D 2>&1 | grep -E "SHARDED_OK|Error|error|assert" >> $L
done; done
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
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$G --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $G --steps 20 --warmup 5 > $O/r02_bench_quick_${G}gpu.json 2> $O/r02_bench_quick_${G}gpu.err
echo "bench N=$G rc=$?"; tail -c 300 $O/r02_bench_quick_${G}gpu.err; show $O/r02_bench_quick_${G}gpu.json
# the shard size of the 8-GPU C4 run (1.25 M customers per GPU) on G GPUs, 200 timed sweeps, both transports
for C in p2p nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$G --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus $G --steps 200 --warmup 10 \
     --customers $((1250000 * G)) --collective $C --no-ess --no-configs --no-forecast --no-cpu-baseline --no-peaks > $O/r02_bench_shard8_${G}gpu_$C.json 2> $O/r02_bench_shard8_${G}gpu_$C.err
  echo "shard8 $C rc=$?"; show $O/r02_bench_shard8_${G}gpu_$C.json
done
