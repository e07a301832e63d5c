#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -q ) > $O/r02_pytest13.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest13.log
tail -12 $O/r02_pytest13.log
