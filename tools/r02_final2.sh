#!/bin/bash
# round-2 check of the committed state on one GPU (after the ragged-size test): full test suite, smoke(), the driver's bench
# command, stage timing of the set-up calls
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -q ) > $O/r02_pytest_full2.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_full2.log
tail -8 $O/r02_pytest_full2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke2.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r02_smoke2.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $O/r02_bench_1gpu_b.json 2> $O/r02_bench_1gpu_b.err; echo "bench rc=$?"
tail -c 300 $O/r02_bench_1gpu_b.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_1gpu_b.json").read().strip().splitlines()[-1])
print("value %.4g ms/step %.4f launches %d e2e %.4g (%s) pageable %.4g stationary %.4g digest ok %s" % (d["value"], d["ms_per_step"], d["gpu_launches"], d["e2e"]["value"], d["e2e"]["seconds_all_runs"], d["e2e_pageable"]["value"], d["stationary"]["value"], d["digest"]["matches_committed"]))
print("roofline", json.dumps(d["roofline"])[:700])
print("forecast", json.dumps(d["forecast"]["roofline"]), d["forecast"]["kernel_ms"])
print("configs", {k: (round(v["wall_s"], 3), round(v["ess_per_sec"], 1)) for k, v in d["configs"].items()})
PY
timeout 200 python tools/e2e_stages.py > $O/r02_e2e_stages.log 2>&1; cat $O/r02_e2e_stages.log
