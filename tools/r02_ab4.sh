#!/bin/bash
# round-2 A/B #4: forecast main pass register-fed vs TMA-fed (both with the global deferred list), parity, posterior, ncu
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
{
for k in reg tma; do
  CLV_FC_KERNEL=$k timeout 120 python tools/forecast_ab.py 23570 800 3
  CLV_FC_KERNEL=$k timeout 120 python tools/forecast_ab.py 1000000 400 3
  CLV_FC_KERNEL=$k timeout 180 python tools/forecast_ab.py 1000000 2000 3
done
} > $O/r02_ab4.log 2>&1
cat $O/r02_ab4.log
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/r02_pytest4.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest4.log
tail -6 $O/r02_pytest4.log
timeout 500 python -m pytest tests/test_gpu_posterior.py -m gpu -q > $O/r02_pytest_post.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_post.log
tail -6 $O/r02_pytest_post.log
# ncu: full set of the sweep kernel and of both forecast main passes (same commands ran above / run first without ncu)
timeout 200 python tools/kernel_ab.py 4000000 24 1 10 fast truth > $O/r02_plain_sweep.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 14 -c 2 -f -o $O/r02_sweep \
    python tools/kernel_ab.py 4000000 24 1 10 fast truth > $O/r02_ncu_sweep.log 2>&1
echo "sweep ncu rc=$?"; tail -1 $O/r02_plain_sweep.log
for k in reg tma; do
  CLV_FC_KERNEL=$k timeout 120 python tools/forecast_ab.py 1000000 200 1 > $O/r02_plain_fc_$k.log 2>&1 &&
  CLV_FC_KERNEL=$k timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_forecast -c 2 -f -o $O/r02_forecast_$k \
      python tools/forecast_ab.py 1000000 200 1 > $O/r02_ncu_fc_$k.log 2>&1
  echo "forecast $k ncu rc=$?"; tail -1 $O/r02_plain_fc_$k.log
done
