#!/bin/bash
# round-2 A/B #7: three / four customers per thread
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
{
for n in 10000000 1250000; do
  CLV_NO_TIMING=1 CLV_SWEEP_MODE=stream timeout 200 python tools/kernel_ab.py $n 200 1 20 fast truth
  for lib in libcpt3_4 libcpt3_3 libcpt4_3; do
    for b in 16 24 32; do
      CLV_B200_LIB=$PWD/build_ab/$lib.so CLV_SWEEP_BLOCKS_PER_SM2=$b CLV_NO_TIMING=1 CLV_SWEEP_MODE=stream timeout 200 python tools/kernel_ab.py $n 200 1 20 fast truth
    done
  done
done
} > $O/r02_ab7.log 2>&1
cut -c1-200 $O/r02_ab7.log
