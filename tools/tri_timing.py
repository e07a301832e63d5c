#!/usr/bin/env python
"""Trivariate vs bivariate sweep time on the same synthetic shard."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcmc_clv_model_b200 import Sampler
from mcmc_clv_model_b200.synthetic import C4_BETA, C4_GAMMA, C4_SEED, C4_T_CAL, generate_cbs_arrays
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
c = generate_cbs_arrays(n, C4_BETA, C4_GAMMA, T_cal=C4_T_CAL, seed=C4_SEED, with_truth=False)
log_s = 3.2 + 0.3 * c["X"][:, 1] + 0.6 * np.cos(np.arange(n) * 0.7)
for D in (2, 3):
    with Sampler(c["x"], c["t_x"], c["T_cal"], c["X"], log_s if D == 3 else None, model_dim=D, chains=1, seed=42, sweep_mode="stream") as s:
        s.advance(20)
        ms = s.advance_timed(40)
    print(f"D={D}: {ms/40:.4f} ms/sweep -> {n*40/(ms*1e-3):.4g} customer-updates/s")
