#!/bin/bash
# round-2 #14: digest check of the pre-scaled statistics, tail tuning of the dynamic tiles with the E32 kernel, ncu capture + launch list
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/r02_pytest14.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest14.log
tail -4 $O/r02_pytest14.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-ess --no-configs --no-forecast --no-cpu-baseline --no-peaks > $O/r02_bench14.json 2> $O/r02_bench14.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench14.json").read().strip().splitlines()[-1])
print("value %.4g ms/step %.4f e2e %.4g digest %s %s ok %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["digest"]["level_2_sha256"][:16], d["digest"]["level_1_hash64"], d["digest"]["matches_committed"]))
PY
{
for lib in mcmc_clv_model_b200/libclv_b200.so build_ab/libprev.so; do
  for n in 10000000 1250000; do
    CLV_B200_LIB=$PWD/$lib CLV_NO_TIMING=1 CLV_SWEEP_MODE=stream timeout 200 python tools/kernel_ab.py $n 300 1 20 fast truth
  done
done
for r in 0 2 3; do
  CLV_SWEEP_SMALL_ROUNDS=$r CLV_NO_TIMING=1 CLV_SWEEP_MODE=stream timeout 200 python tools/kernel_ab.py 1250000 300 1 20 fast truth
done
CLV_SWEEP_DYNAMIC=0 CLV_NO_TIMING=1 CLV_SWEEP_MODE=stream timeout 200 python tools/kernel_ab.py 1250000 300 1 20 fast truth
CLV_SWEEP_DYNAMIC=0 CLV_NO_TIMING=1 CLV_SWEEP_MODE=stream timeout 200 python tools/kernel_ab.py 10000000 300 1 20 fast truth
for n in 2500000 5000000; do
  CLV_NO_TIMING=1 CLV_SWEEP_MODE=stream timeout 200 python tools/kernel_ab.py $n 300 1 20 fast truth
done
timeout 200 python tools/small_n_timing.py 4 3000 abe 2 | grep -i "stream\|persist"
} > $O/r02_ab14.log 2>&1
cut -c1-200 $O/r02_ab14.log
python tools/kernel_ab.py 4000000 24 1 10 fast truth > $O/r02_plain_sweep.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 14 -c 2 -f -o $O/r02_sweep5 \
    python tools/kernel_ab.py 4000000 24 1 10 fast truth > $O/r02_ncu_sweep.log 2>&1
echo "sweep ncu rc=$?"; tail -2 $O/r02_plain_sweep.log
timeout 300 python bench.py --steps 4 --warmup 3 --no-ess --no-configs --no-forecast --no-cpu-baseline --no-peaks > $O/r02_plain_bench.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-ess --no-configs --no-forecast --no-cpu-baseline --no-peaks > $O/r02_ncu_bench.log 2>&1
echo "launch list rc=$?"; wc -l $O/r02_launches.csv
