#!/bin/bash
# round-2 A/B #16: variates of VAR_AHEAD Metropolis steps drawn side by side in the one-customer kernels (latency-bound regime)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/r02_pytest16.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest16.log
tail -4 $O/r02_pytest16.log
{
for lib in mcmc_clv_model_b200/libclv_b200.so build_ab/libprev.so; do
  CLV_B200_LIB=$PWD/$lib timeout 200 python tools/small_n_timing.py 4 3000 abe 2 | grep -i "stream\|persist"
  CLV_B200_LIB=$PWD/$lib timeout 200 python tools/small_n_timing.py 2 3000 full 2 | grep -i "stream\|persist"
  CLV_B200_LIB=$PWD/$lib timeout 200 python tools/small_n_timing.py 2 3000 full 3 | grep -i "stream\|persist"
  for n in 10000000 1250000; do
    CLV_B200_LIB=$PWD/$lib CLV_NO_TIMING=1 CLV_SWEEP_MODE=stream timeout 200 python tools/kernel_ab.py $n 300 1 20 fast truth
  done
  CLV_B200_LIB=$PWD/$lib CLV_SWEEP_SMALL_ROUNDS=2 CLV_NO_TIMING=1 CLV_SWEEP_MODE=stream timeout 200 python tools/kernel_ab.py 1250000 300 1 20 fast truth
done
} > $O/r02_ab16.log 2>&1
cut -c1-200 $O/r02_ab16.log
