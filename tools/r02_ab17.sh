#!/bin/bash
# round-2 A/B #17: rounds of 128-customer tiles at the end of a shard (1 vs 2) over shard sizes
cd "$(dirname "$0")/.."
O=gpurun_out
{
for n in 400000 625000 1250000 2500000 5000000 10000000; do
  for r in 1 2; do
    CLV_SWEEP_SMALL_ROUNDS=$r CLV_NO_TIMING=1 CLV_SWEEP_MODE=stream timeout 200 python tools/kernel_ab.py $n 300 1 20 fast truth | sed "s/^/rounds=$r /"
  done
done
} > $O/r02_ab17.log 2>&1
cut -c1-175 $O/r02_ab17.log
