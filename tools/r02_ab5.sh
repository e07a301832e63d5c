#!/bin/bash
# round-2 A/B #5: sweep kernel with two customers per thread (CLV_SWEEP_CPT=2) vs one
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
CLV_SWEEP_CPT=2 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "trajectory or lockstep or odd_step or many_cov or chain_offset or checkpoint or persistent" > $O/r02_pytest_cpt2.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_cpt2.log
tail -4 $O/r02_pytest_cpt2.log
{
for n in 10000000 1250000; do
  CLV_NO_TIMING=1 CLV_SWEEP_MODE=stream timeout 200 python tools/kernel_ab.py $n 200 1 20 fast truth
  for lib in mcmc_clv_model_b200/libclv_b200.so build_ab/libcpt2_4.so build_ab/libcpt2_6.so; do
    CLV_SWEEP_CPT=2 CLV_B200_LIB=$PWD/$lib CLV_NO_TIMING=1 CLV_SWEEP_MODE=stream timeout 200 python tools/kernel_ab.py $n 200 1 20 fast truth
  done
  for b in 10 20 30; do
    CLV_SWEEP_CPT=2 CLV_SWEEP_BLOCKS_PER_SM2=$b CLV_NO_TIMING=1 CLV_SWEEP_MODE=stream timeout 200 python tools/kernel_ab.py $n 200 1 20 fast truth
  done
done
CLV_SWEEP_CPT=2 timeout 200 python tools/small_n_timing.py 4 6000 abe 2 | grep stream
timeout 200 python tools/small_n_timing.py 4 6000 abe 2 | grep stream
CLV_SWEEP_CPT=2 timeout 200 python tools/small_n_timing.py 56 3000 abe 2 | grep stream
timeout 200 python tools/small_n_timing.py 56 3000 abe 2 | grep stream
CLV_SWEEP_CPT=2 timeout 200 python tools/small_n_timing.py 2 3000 full 2 | grep stream
timeout 200 python tools/small_n_timing.py 2 3000 full 2 | grep stream
} > $O/r02_ab5.log 2>&1
cat $O/r02_ab5.log | cut -c1-220
