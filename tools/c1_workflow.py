#!/usr/bin/env python
"""The reference's own workflow (run_mcmc_abe.py:61-95 + analysis_abe.py:405) through the drop-in modules, timed."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd
from src.models.bivariate.mcmc import mcmc_draw_parameters, draw_future_transactions
from mcmc_clv_model_b200.analysis import table4_inputs
from mcmc_clv_model_b200.diagnostics import summarize
d = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "cdnow_abe.npz"))
cbs = pd.DataFrame({k: d[k] for k in ("x", "t_x", "T_cal", "first_sales_scaled", "x_star")})
mcmc_draw_parameters(cbs, mcmc=10, burnin=10, thin=1, chains=1, seed=1, trace=0)     # context warm-up
for name, cov in (("M1", []), ("M2", ["first_sales_scaled"])):
    t0 = time.perf_counter()
    draws = mcmc_draw_parameters(cal_cbs=cbs, covariates=cov, mcmc=4000, burnin=10000, thin=1, chains=4, seed=42, trace=0, n_mh_steps=20)
    t1 = time.perf_counter()
    xs = draw_future_transactions(cbs, draws, T_star=39.0, seed=42)
    t2 = time.perf_counter()
    t4 = table4_inputs(draws)
    t3 = time.perf_counter()
    s = summarize(draws["level_2"])
    print(f"{name}: mcmc_draw_parameters {t1-t0:.2f} s | draw_future_transactions {t2-t1:.2f} s | table-4 inputs {t3-t2:.2f} s | "
          f"level_2 means {[round(s[j]['mean'], 3) for j in s]} | mean x* {xs.mean():.3f} (hold-out actual {cbs['x_star'].mean():.3f}) | "
          f"corr(E[x*], x*) {np.corrcoef(xs.mean(axis=0), cbs['x_star'])[0,1]:.3f} | P(alive) {t4['P(alive at T_cal)'].mean():.3f}", flush=True)
