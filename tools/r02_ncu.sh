#!/bin/bash
# round-2 ncu captures (one GPU): full set of the sweep kernel and of the forecast kernel, launch list of a short bench
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python tools/kernel_ab.py 4000000 24 1 10 fast truth > $O/r02_plain_sweep.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 14 -c 2 -f -o $O/r02_sweep \
    python tools/kernel_ab.py 4000000 24 1 10 fast truth > $O/r02_ncu_sweep.log 2>&1
echo "sweep ncu rc=$?"; tail -2 $O/r02_plain_sweep.log
python tools/forecast_ab.py 1000000 200 2 > $O/r02_plain_fc.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_forecast -c 4 -f -o $O/r02_forecast \
    python tools/forecast_ab.py 1000000 200 2 > $O/r02_ncu_fc.log 2>&1
echo "forecast ncu rc=$?"; tail -2 $O/r02_plain_fc.log
python bench.py --steps 4 --warmup 3 --no-ess --no-configs --no-forecast --no-cpu-baseline --no-peaks > $O/r02_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-ess --no-configs --no-forecast --no-cpu-baseline --no-peaks > $O/r02_ncu_bench.log 2>&1
echo "launch list rc=$?"; wc -l $O/r02_launches.csv
