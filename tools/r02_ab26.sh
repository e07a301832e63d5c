#!/bin/bash
# round-2 A/B #26: one vs two customers per thread in the many-chains regime of C1 (chains x 2 357 customers), E32 kernels
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
{
for ch in 56 42 84 112; do
  for cpt in 1 2; do
    CLV_SWEEP_CPT=$cpt timeout 200 python tools/small_n_timing.py $ch 3000 abe 2 | grep "mode=stream" | sed "s/^/cpt=$cpt /"
  done
done
} > $O/r02_ab26.log 2>&1
cat $O/r02_ab26.log
