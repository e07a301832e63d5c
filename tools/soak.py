#!/usr/bin/env python
"""Soak: long persistent-kernel runs (grid barrier, rotating accumulators) on the CDNOW data, bivariate and trivariate."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcmc_clv_model_b200 import Sampler
d = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "cdnow_abe.npz"))
X = np.column_stack([np.ones(d["x"].size), d["first_sales_scaled"]])
for D, chains, sweeps in ((2, 4, 60000), (2, 64, 20000), (3, 16, 30000), (2, 300, 5000)):
    t0 = time.perf_counter()
    with Sampler(d["x"], d["t_x"], d["T_cal"], X, d["log_s"] if D == 3 else None, model_dim=D, chains=chains, seed=1) as s:
        out = s.run(sweeps - 100, 100, 10, store_level1=False)
    dt = time.perf_counter() - t0
    l2 = out["level_2"]
    assert np.isfinite(l2).all()
    print(f"D={D} chains={chains} sweeps={sweeps}: {dt:.2f} s ({dt/sweeps*1e6:.1f} us/sweep), mean level_2 = {np.round(l2.mean(axis=(0,1)), 3)}", flush=True)
