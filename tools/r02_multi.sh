#!/bin/bash
# round-2 multi-GPU validation (gpurun --gpus G): sharded == 1-GPU chain over both transports, multi-GPU tests, bench digests
cd "$(dirname "$0")/.."
G=${1:-2}
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/r02_pytest_multi_${G}gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_multi_${G}gpu.log
tail -3 $O/r02_pytest_multi_${G}gpu.log
L=$O/r02_sharded_check_${G}gpu.log
: > $L
for D in 2 3; do for P in 1 0; do
  CLV_P2P=$P timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$G --master-addr 127.0.0.1 --master-port 29551 \
     tools/sharded_check.py 1000003 $D 2>&1 | grep -E "SHARDED_OK|Error|error|assert" >> $L
done; done
cat $L
for N in 1 $G; do
  if [ $N = 1 ]; then python bench.py --gpus 1 --steps 20 --warmup 5 --no-configs --no-forecast --no-cpu-baseline > $O/r02_bench_quick_1gpu.json 2> $O/r02_bench_quick_1gpu.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 20 --warmup 5 > $O/r02_bench_quick_${N}gpu.json 2> $O/r02_bench_quick_${N}gpu.err; fi
  echo "bench N=$N rc=$?"; tail -c 600 $O/r02_bench_quick_${N}gpu.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/r02_bench_quick_${N}gpu.json").read().strip().splitlines()[-1])
    print("N=$N value %.4g ms/step %.4f e2e %.4g digest %s %s launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["digest"]["level_2_sha256"][:16], d["digest"]["level_1_hash64"], d["gpu_launches"]))
    print(" stationary %.4g  ess %s" % (d["stationary"]["value"], json.dumps(d["ess"])[:600]))
except Exception as e:
    print("no JSON line:", e)
PY
done
# forecast kernel A/B (1 GPU of the box)
for lib in mcmc_clv_model_b200/libclv_b200.so build_ab/libfc5.so build_ab/libfc4.so; do
  CLV_B200_LIB=$PWD/$lib python tools/forecast_ab.py 1000000 400 3 2>&1 | tail -1
done
