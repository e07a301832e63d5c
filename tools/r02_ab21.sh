#!/bin/bash
# round-2 A/B #21: host side of the staged draw transfers: copy threads x ring slot size (12 GB of level-1 draws, C2 shape x 4 chains)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
{
nproc; grep -m1 "model name" /proc/cpuinfo
for cfg in "8 16" "4 16" "12 16" "16 16" "8 4" "8 64" "16 64" "2 16"; do set -- $cfg
  echo "== CLV_COPY_THREADS=$1 CLV_STAGE_PIECE_MB=$2"
  CLV_COPY_THREADS=$1 CLV_STAGE_PIECE_MB=$2 timeout 300 python tools/c2_timing.py 2>&1 | grep "store_level1=True"
done
} > $O/r02_ab21.log 2>&1
cat $O/r02_ab21.log
