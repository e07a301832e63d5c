#!/bin/bash
# last check of the committed library on one GPU: the parity suite and smoke()
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x ) > $O/r02_pytest_final4.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_final4.log
tail -5 $O/r02_pytest_final4.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
