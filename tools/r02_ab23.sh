#!/bin/bash
# round-2 #23: what bounds the host side of the level-1 transfer (fresh / touched / huge-page destination)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python tools/d2h_probe.py > $O/r02_d2h_probe.log 2>&1
cat $O/r02_d2h_probe.log
free -g | head -2
