#!/bin/bash
# round-2 A/B #1: level-2 rewrite + PDL vs the round-1 library
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r02_pytest1.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest1.log
tail -5 $O/r02_pytest1.log
{
for lib in build_ab/libr01.so mcmc_clv_model_b200/libclv_b200.so; do
  CLV_B200_LIB=$PWD/$lib python tools/small_n_timing.py 4 6000 abe 2
  CLV_B200_LIB=$PWD/$lib python tools/small_n_timing.py 2 3000 full 2
  CLV_B200_LIB=$PWD/$lib python tools/small_n_timing.py 2 3000 full 3
done
for n in 1250000 10000000; do
  CLV_NO_TIMING=1 CLV_SWEEP_MODE=stream CLV_B200_LIB=$PWD/build_ab/libr01.so python tools/kernel_ab.py $n 200 1 20 fast truth
  CLV_NO_TIMING=1 CLV_SWEEP_MODE=stream CLV_NO_PDL=1 python tools/kernel_ab.py $n 200 1 20 fast truth
  CLV_NO_TIMING=1 CLV_SWEEP_MODE=stream python tools/kernel_ab.py $n 200 1 20 fast truth
  CLV_NO_TIMING=1 CLV_SWEEP_MODE=stream CLV_B200_LIB=$PWD/build_ab/libexp8.so python tools/kernel_ab.py $n 200 1 20 fast truth
done
for b in 8 12 16 24 32 64; do
  CLV_SWEEP_BLOCKS_PER_SM=$b CLV_NO_TIMING=1 CLV_SWEEP_MODE=stream python tools/kernel_ab.py 1250000 200 1 20 fast truth
done
python tools/kernel_ab.py 1250000 200 1 20 fast truth
python tools/kernel_ab.py 10000000 100 1 20 fast truth
} > $O/r02_ab1.log 2>&1
cat $O/r02_ab1.log | grep -v "^$"
