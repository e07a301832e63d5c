#!/bin/bash
# round-2, 8 GPUs: sharded == 1-GPU chain over both transports, bench at 8 (p2p and NCCL) and 4 GPUs
cd "$(dirname "$0")/.."
G=${1:-8}
O=gpurun_out
mkdir -p $O
L=$O/r02_sharded_check_${G}gpu.log
: > $L
for D in 2 3; do for P in 1 0; do
  CLV_P2P=$P timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$G --master-addr 127.0.0.1 --master-port 29561 \
     tools/sharded_check.py 2000003 $D 2>&1 | grep -E "SHARDED_OK|Error|error|assert" >> $L
done; done
cat $L
run_bench() {  # N collective tag
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$1 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $1 --steps 20 --warmup 5 --collective $2 > $O/r02_bench_$3.json 2> $O/r02_bench_$3.err
  echo "bench $3 rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("$O/r02_bench_$3.json").read().strip().splitlines()[-1])
    print("$3: value %.4g ms/step %.4f e2e %.4g stationary %.4g digest %s %s match %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["stationary"]["value"], d["digest"]["level_2_sha256"][:16], d["digest"]["level_1_hash64"], d["digest"].get("matches_committed")))
    print("   ess strong %s weak %s" % (json.dumps(d["ess"]["strong"])[:300], json.dumps(d["ess"].get("weak"))[:300]))
except Exception as e:
    print("no JSON line:", e)
PY
}
run_bench $G p2p ${G}gpu
run_bench $G nccl ${G}gpu_nccl
run_bench 4 p2p 4gpu
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$G --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus $G --steps 200 --warmup 10 --no-ess > $O/r02_bench_${G}gpu_200.json 2> $O/r02_bench_${G}gpu_200.err
python -c "
import json
d=json.loads(open('$O/r02_bench_${G}gpu_200.json').read().strip().splitlines()[-1]); print('200 steps: value %.4g ms/step %.4f' % (d['value'], d['ms_per_step']))"
