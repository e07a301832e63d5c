#!/bin/bash
# round-2 A/B #2: PDL trigger placement, persistent vs stream at a 1.25 M shard, lockstep / FAST-replay tests
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/r02_pytest2.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest2.log
tail -5 $O/r02_pytest2.log
{
for mode in 0 1 2 3; do
  export CLV_PDL_MODE=$mode
  if [ $mode = 0 ]; then export CLV_NO_PDL=1; else unset CLV_NO_PDL; fi
  echo "== CLV_PDL_MODE=$mode (0 = no PDL)"
  python tools/small_n_timing.py 4 6000 abe 2 | grep stream
  python tools/small_n_timing.py 2 3000 full 2 | grep stream
  python tools/small_n_timing.py 2 3000 full 3 | grep stream
  python tools/small_n_timing.py 56 3000 abe 2 | grep stream
  CLV_NO_TIMING=1 CLV_SWEEP_MODE=stream python tools/kernel_ab.py 1250000 200 1 20 fast truth
  CLV_NO_TIMING=1 CLV_SWEEP_MODE=stream CLV_SWEEP_BLOCKS_PER_SM=24 python tools/kernel_ab.py 1250000 200 1 20 fast truth
  CLV_NO_TIMING=1 CLV_SWEEP_MODE=stream python tools/kernel_ab.py 10000000 100 1 20 fast truth
done
unset CLV_PDL_MODE CLV_NO_PDL
echo "== persistent"
python tools/small_n_timing.py 56 3000 abe 2
python tools/small_n_timing.py 8 3000 abe 2
python tools/small_n_timing.py 2 3000 full 2
CLV_NO_TIMING=1 CLV_SWEEP_MODE=persistent python tools/kernel_ab.py 1250000 200 1 20 fast truth
CLV_NO_TIMING=1 CLV_SWEEP_MODE=persistent python tools/kernel_ab.py 10000000 100 1 20 fast truth
} > $O/r02_ab2.log 2>&1
grep -v "^$" $O/r02_ab2.log
