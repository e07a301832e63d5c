#!/bin/bash
# round-2 A/B #12: full GPU suite on the E32 kernel, one-customer kernel at 8 vs 4 blocks per SM in the small-N regime, ncu capture
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > $O/r02_pytest12.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest12.log
tail -8 $O/r02_pytest12.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r02_smoke.log
{
for lib in mcmc_clv_model_b200/libclv_b200.so build_ab/libk1_mb4.so; do
  CLV_B200_LIB=$PWD/$lib timeout 200 python tools/small_n_timing.py 4 3000 abe 2 | grep -i "stream\|persist"
  CLV_B200_LIB=$PWD/$lib timeout 200 python tools/small_n_timing.py 2 3000 full 2 | grep -i "stream\|persist"
  CLV_B200_LIB=$PWD/$lib timeout 200 python tools/small_n_timing.py 2 3000 full 3 | grep -i "stream\|persist"
  CLV_B200_LIB=$PWD/$lib CLV_SWEEP_CPT=1 CLV_NO_TIMING=1 CLV_SWEEP_MODE=stream timeout 200 python tools/kernel_ab.py 1250000 300 1 20 fast truth
done
} > $O/r02_ab12.log 2>&1
cut -c1-200 $O/r02_ab12.log
python tools/kernel_ab.py 4000000 24 1 10 fast truth > $O/r02_plain_sweep.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 14 -c 2 -f -o $O/r02_sweep4 \
    python tools/kernel_ab.py 4000000 24 1 10 fast truth > $O/r02_ncu_sweep.log 2>&1
echo "sweep ncu rc=$?"; tail -2 $O/r02_plain_sweep.log
