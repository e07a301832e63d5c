#!/bin/bash
# round-2 A/B #8: dynamic tiles with a finer tail for the two-customer sweep kernel
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/r02_pytest8.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest8.log
tail -5 $O/r02_pytest8.log
{
for n in 1250000 10000000 400000; do
  CLV_SWEEP_DYNAMIC=0 CLV_NO_TIMING=1 CLV_SWEEP_MODE=stream timeout 200 python tools/kernel_ab.py $n 300 1 20 fast truth
  for r in 0 1 2 3 4; do
    CLV_SWEEP_SMALL_ROUNDS=$r CLV_NO_TIMING=1 CLV_SWEEP_MODE=stream timeout 200 python tools/kernel_ab.py $n 300 1 20 fast truth
  done
done
timeout 200 python tools/small_n_timing.py 56 3000 abe 2 | grep stream
CLV_SWEEP_DYNAMIC=0 timeout 200 python tools/small_n_timing.py 56 3000 abe 2 | grep stream
} > $O/r02_ab8.log 2>&1
cut -c1-200 $O/r02_ab8.log
