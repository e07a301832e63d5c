#!/bin/bash
# round-2 A/B #6: grid of the two-customer sweep kernel; parity suite on the new defaults; ncu of k_sweep2
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
{
for b in 12 16 20 24; do CLV_SWEEP_BLOCKS_PER_SM2=$b CLV_NO_TIMING=1 CLV_SWEEP_MODE=stream timeout 200 python tools/kernel_ab.py 1250000 300 1 20 fast truth; done
for b in 24 30 40 60; do CLV_SWEEP_BLOCKS_PER_SM2=$b CLV_NO_TIMING=1 CLV_SWEEP_MODE=stream timeout 200 python tools/kernel_ab.py 10000000 100 1 20 fast truth; done
CLV_NO_TIMING=1 CLV_SWEEP_MODE=stream timeout 200 python tools/kernel_ab.py 10000000 100 1 20 fast reference
CLV_SWEEP_CPT=1 CLV_NO_TIMING=1 CLV_SWEEP_MODE=stream timeout 200 python tools/kernel_ab.py 10000000 100 1 20 fast reference
} > $O/r02_ab6.log 2>&1
cut -c1-200 $O/r02_ab6.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/r02_pytest6.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest6.log
tail -5 $O/r02_pytest6.log
timeout 200 python tools/kernel_ab.py 4000000 24 1 10 fast truth > $O/r02_plain_sweep.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 14 -c 2 -f -o $O/r02_sweep2 \
    python tools/kernel_ab.py 4000000 24 1 10 fast truth > $O/r02_ncu_sweep.log 2>&1
echo "sweep ncu rc=$?"; tail -1 $O/r02_plain_sweep.log
