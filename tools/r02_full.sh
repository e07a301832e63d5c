#!/bin/bash
# full GPU test suite + the driver's bench commands
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( time python -m pytest tests -m gpu -q ) > $O/r02_pytest_full.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_full.log
tail -25 $O/r02_pytest_full.log
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > $O/r02_bench_1gpu.json 2> $O/r02_bench_1gpu.err; echo "bench rc=$?"
tail -c 1500 $O/r02_bench_1gpu.err
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > $O/r02_bench_ref.json 2> $O/r02_bench_ref.err; echo "ref rc=$?"
tail -c 400 $O/r02_bench_ref.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_1gpu.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "gpu_launches"): print(k, d[k])
print("e2e", d["e2e"]["value"], "pageable", d["e2e_pageable"]["value"], "stationary", d["stationary"]["value"])
print("digest", d["digest"])
print("cpu", d["cpu_baseline"])
print("ess", json.dumps(d["ess"])[:1500])
print("configs", json.dumps(d["configs"])[:3000])
print("forecast", json.dumps(d["forecast"])[:2000])
r = json.loads(open("gpurun_out/r02_bench_ref.json").read().strip().splitlines()[-1])
print("ref", r["value"], r["cpu_baseline"], r.get("cpu_port"))
PY
