#!/bin/bash
# round-2 A/B #19: resident forecast in chunks of draw pairs, the second pass of a chunk on a side stream beside the main pass
# of the next (CLV_FC_CHUNKS x CLV_FC_SIDE_BLOCKS); forecast tests first
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "forecast" > $O/r02_pytest19.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest19.log
tail -4 $O/r02_pytest19.log
{
for nd in 400 2000; do
  for cfg in "1 1" "4 1" "8 1" "8 2" "16 1" "16 2" "32 1"; do set -- $cfg
    CLV_FC_CHUNKS=$1 CLV_FC_SIDE_BLOCKS=$2 timeout 300 python tools/forecast_ab.py 1000000 $nd 4 | sed "s/^/chunks=$1 side=$2 /"
  done
done
} > $O/r02_ab19.log 2>&1
cat $O/r02_ab19.log
