#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
{
for lib in mcmc_clv_model_b200/libclv_b200.so build_ab/libafloat.so; do
  for n in 10000000 1250000; do
    CLV_B200_LIB=$PWD/$lib CLV_NO_TIMING=1 CLV_SWEEP_MODE=stream timeout 200 python tools/kernel_ab.py $n 300 1 20 fast truth
  done
done
} > $O/r02_ab15.log 2>&1
cut -c1-200 $O/r02_ab15.log
