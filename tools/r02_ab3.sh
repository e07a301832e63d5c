#!/bin/bash
# round-2 A/B #3: TMA-fed forecast kernel vs the register-fed one; posterior tests; forecast parity tests
# (every step under its own timeout: a kernel that waits on an mbarrier must not hold the box)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
{
for k in tma reg; do
  CLV_FC_KERNEL=$k timeout 120 python tools/forecast_ab.py 23570 800 3
  CLV_FC_KERNEL=$k timeout 120 python tools/forecast_ab.py 1000000 400 3
  CLV_FC_KERNEL=$k timeout 180 python tools/forecast_ab.py 1000000 2000 3
done
} > $O/r02_ab3.log 2>&1
cat $O/r02_ab3.log
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/r02_pytest3.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest3.log
tail -6 $O/r02_pytest3.log
timeout 400 python -m pytest tests/test_gpu_posterior.py -m gpu -q > $O/r02_pytest_post.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_post.log
tail -12 $O/r02_pytest_post.log
