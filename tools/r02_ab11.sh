#!/bin/bash
# round-2 A/B #11: batched input loads, covariates kept in shared memory, one barrier per dynamic tile
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/r02_pytest11.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest11.log
tail -5 $O/r02_pytest11.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-ess --no-configs --no-forecast --no-cpu-baseline --no-peaks > $O/r02_bench11.json 2> $O/r02_bench11.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench11.json").read().strip().splitlines()[-1])
print("value %.4g ms/step %.4f e2e %.4g digest %s %s ok %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["digest"]["level_2_sha256"][:16], d["digest"]["level_1_hash64"], d["digest"]["matches_committed"]))
PY
{
for lib in mcmc_clv_model_b200/libclv_b200.so $EXTRA_LIBS; do
  for n in 10000000 1250000; do
    CLV_B200_LIB=$PWD/$lib CLV_NO_TIMING=1 CLV_SWEEP_MODE=stream timeout 200 python tools/kernel_ab.py $n 300 1 20 fast truth
  done
  CLV_B200_LIB=$PWD/$lib timeout 200 python tools/small_n_timing.py 56 3000 abe 2 | grep -i "stream\|persist"
done
} > $O/r02_ab11.log 2>&1
cut -c1-220 $O/r02_ab11.log
