#!/usr/bin/env python
"""The reference's full-data run (C2: 23 570 customers, K=2, 4 chains x (10 000 burn-in + 4 000 kept, thin 1),
run_mcmc_full.py) through Sampler.run: how much of the wall time is the device->host transfer of the 12 GB of level-1
draws?  Usage: python tools/c2_timing.py [mcmc]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcmc_clv_model_b200 import Sampler

mcmc = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
d = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "cdnow_full.npz"))
n = d["x"].size
X = np.column_stack([np.ones(n), d["first_sales_scaled"]])
args = (d["x"].astype(np.int32), d["t_x"], d["T_cal"], X)
for store in (False, True, True):
    t0 = time.perf_counter()
    with Sampler(*args, chains=4, seed=42) as s:
        out = s.run(10000, mcmc, 1, store_level1=store)
    dt = time.perf_counter() - t0
    gb = 0.0 if out["level_1"] is None else out["level_1"].nbytes / 1e9
    print(f"store_level1={store}: {dt:.2f} s wall, {gb:.1f} GB of level-1 draws to the host"
          + (f" ({gb / dt:.1f} GB/s if the transfer were all of it)" if gb else ""), flush=True)
    del out
