#!/usr/bin/env python
"""Small end-to-end exercise of every kernel, for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pandas as pd
from mcmc_clv_model_b200 import Sampler
from mcmc_clv_model_b200.cbs import elog2cbs
from mcmc_clv_model_b200.synthetic import C4_BETA, C4_GAMMA, generate_cbs_arrays
from mcmc_clv_model_b200.api import _forecast

g = generate_cbs_arrays(777, C4_BETA, C4_GAMMA, T_cal=(27.0, 38.9), seed=5)
rng = np.random.default_rng(0)
log_s = rng.normal(3.0, 0.6, 777)
for D in (2, 3):
    for mode in ("stream", "persistent"):
        with Sampler(g["x"], g["t_x"], g["T_cal"], g["X"], log_s if D == 3 else None, model_dim=D, chains=3, seed=1,
                     sweep_mode=mode, n_mh_steps=5) as s:
            out = s.run(3, 5, 2)
            s.advance(2)
            fr = s.run(0, 3, 1)
            f = s.forecast_resident(seed=3, want_x_star=True)
            summ = s.posterior_summary()
            wk = s.weekly_tracking(rng.uniform(0, 5, 777), np.arange(1.0, 20.0), seed=4)
            assert np.isfinite(out["level_2"]).all() and np.isfinite(summ["mean_lambda"]).all() and np.isfinite(wk).all()
    x, sp = _forecast(g["T_cal"], list(out["level_1"]), 39.0, 5, D == 3, 0.5)
with Sampler(g["x"], g["t_x"], g["T_cal"], g["X"], chains=1, seed=1, rng="strict", n_mh_steps=3) as s:
    s.run(1, 2, 1)
elog = pd.DataFrame({"cust": rng.integers(1, 50, 400), "date": pd.Timestamp("2020-01-01") + pd.to_timedelta(rng.integers(0, 300, 400), unit="D"),
                     "sales": rng.random(400)})
cbs = elog2cbs(elog, units="W", T_cal="2020-07-01", T_tot="2020-10-30")
print("sanitize_small ok", len(cbs), x.shape)
