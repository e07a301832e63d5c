#!/bin/bash
# round-2 check of the committed state on one GPU: full test suite, smoke(), the driver's bench commands, launch list
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -q ) > $O/r02_pytest_full3.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_full3.log
tail -8 $O/r02_pytest_full3.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke3.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r02_smoke3.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $O/r02_bench_1gpu_c.json 2> $O/r02_bench_1gpu_c.err; echo "bench rc=$?"
tail -c 300 $O/r02_bench_1gpu_c.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_1gpu_c.json").read().strip().splitlines()[-1])
print("value %.4g ms/step %.4f launches %d e2e %.4g (%s) pageable %.4g stationary %.4g digest ok %s" % (d["value"], d["ms_per_step"], d["gpu_launches"], d["e2e"]["value"], d["e2e"]["seconds_all_runs"], d["e2e_pageable"]["value"], d["stationary"]["value"], d["digest"]["matches_committed"]))
print("roofline", json.dumps(d["roofline"])[:700])
print("forecast", json.dumps(d["forecast"]["roofline"]), d["forecast"]["kernel_ms"])
print("configs", {k: (round(v["wall_s"], 3), round(v["ess_per_sec"], 1)) for k, v in d["configs"].items()})
PY
timeout 300 python bench.py --steps 4 --warmup 3 --no-ess --no-configs --no-forecast --no-cpu-baseline --no-peaks > $O/r02_plain_bench_c.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches_c.csv \
    python bench.py --steps 4 --warmup 3 --no-ess --no-configs --no-forecast --no-cpu-baseline --no-peaks > $O/r02_ncu_bench_c.log 2>&1
echo "launch list rc=$?"; wc -l $O/r02_launches_c.csv
