#!/usr/bin/env python
"""Run on a GPU box: writes gpurun_out/gpu_draws_small.pkl, a small draws dict produced by the CUDA sampler through the
drop-in entry points (for the CPU-side consumer-contract test against the reference's analysis helpers)."""
import os, pickle, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd
from src.models.bivariate.mcmc import mcmc_draw_parameters, draw_future_transactions
d = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "cdnow_abe.npz"))
n = 200
cbs = pd.DataFrame({k: d[k][:n] for k in ("x", "t_x", "T_cal", "first_sales_scaled")})
draws = mcmc_draw_parameters(cbs, covariates=["first_sales_scaled"], mcmc=60, burnin=300, thin=2, chains=2, seed=42, trace=0)
xs = draw_future_transactions(cbs, draws, T_star=39.0, seed=42)
os.makedirs("gpurun_out", exist_ok=True)
pickle.dump(dict(draws=draws, x_star=xs, n=n), open("gpurun_out/gpu_draws_small.pkl", "wb"))
print("wrote gpurun_out/gpu_draws_small.pkl", xs.shape)
