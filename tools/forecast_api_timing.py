#!/usr/bin/env python
"""draw_future_transactions' path with host arrays in and out (clv_forecast): n customers x nd draws, best of three."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcmc_clv_model_b200.api import _forecast
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
nd = int(sys.argv[2]) if len(sys.argv) > 2 else 64
rng = np.random.default_rng(0)
T = rng.uniform(27, 39, n)
l1 = np.empty((nd, n, 4))
l1[..., 0] = rng.lognormal(-3.3, 1.0, (nd, n)); l1[..., 1] = 0.03; l1[..., 2] = T + rng.exponential(30.0, (nd, n)); l1[..., 3] = rng.random((nd, n)) < 0.4
_forecast(T, [l1[:4]], 39.0, 42, False, 0.5)
ts = []
for _ in range(3):
    t0 = time.perf_counter(); xs, _ = _forecast(T, [l1], 39.0, 42, False, 0.5); ts.append(time.perf_counter() - t0); del xs
print(f"clv_forecast n={n} draws={nd}: {[round(t, 3) for t in ts]} s  best {n * nd / min(ts):.4g} cells/s  {n * nd * 40 / min(ts) / 1e9:.1f} GB/s of host traffic")
