#!/usr/bin/env python
"""Per-sweep latency of the small-N regime (C1: CDNOW Abe subset, 2 357 customers): persistent and stream modes.
Usage: [CLV_B200_LIB=...] python tools/small_n_timing.py [chains] [sweeps] [dataset abe|full] [D]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcmc_clv_model_b200 import Sampler

chains = int(sys.argv[1]) if len(sys.argv) > 1 else 4
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
name = sys.argv[3] if len(sys.argv) > 3 else "abe"
D = int(sys.argv[4]) if len(sys.argv) > 4 else 2
d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", f"cdnow_{name}.npz"))
n = d["x"].size
X = np.ones((n, 1)) if name == "abe" and D == 2 else np.column_stack([np.ones(n), d["first_sales_scaled"]])
for mode in ("persistent", "stream"):
    with Sampler(d["x"], d["t_x"], d["T_cal"], X, d["log_s"] if D == 3 else None, model_dim=D, chains=chains, seed=42, sweep_mode=mode) as s:
        s.advance(200)
        ms = s.advance_timed(sweeps)
        st = s.get_state(0)
    print(f"lib={os.path.basename(os.environ.get('CLV_B200_LIB', 'default'))} {name} n={n} D={D} chains={chains} mode={mode}: "
          f"{ms / sweeps * 1e3:.2f} us/sweep  S00={st['Sigma'][0, 0]:.5g}", flush=True)
