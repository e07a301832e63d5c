#!/usr/bin/env python
"""Forecast kernel harness: N synthetic customers x D resident draws; prints the best kernel time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcmc_clv_model_b200 import Sampler
from mcmc_clv_model_b200.synthetic import C4_BETA, C4_GAMMA, C4_SEED, C4_T_CAL, generate_cbs_arrays
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
nd = int(sys.argv[2]) if len(sys.argv) > 2 else 200
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
fc = generate_cbs_arrays(n, C4_BETA, C4_GAMMA, T_cal=C4_T_CAL, T_star=39.0, seed=C4_SEED + 1, with_truth=True)
with Sampler(fc["x"], fc["t_x"], fc["T_cal"], fc["X"], chains=1, seed=7) as s:
    s.set_state(0, log_lambda=np.log(fc["lambda_true"]), log_mu=np.log(fc["mu_true"]), beta=C4_BETA, Sigma=C4_GAMMA)
    s.run_resident(20, nd, 1)
    best = min(s.forecast_resident(T_star=39.0, seed=42)["kernel_ms"] for _ in range(reps))
cells = n * nd
print(os.path.basename(os.environ.get("CLV_B200_LIB","default")), end=" "); print(f"forecast n={n} draws={nd}: {best:.3f} ms  {cells/(best*1e-3):.4g} cells/s  {cells*32/(best*1e-3)/1e9:.0f} GB/s")
