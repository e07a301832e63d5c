"""Device versions of the reductions the reference's analysis scripts run over the draws dict
(src/models/utils/analysis_bi_helpers.py; bivariate/analysis_abe.py:446-464).  They take the same `draws` dict that
`mcmc_draw_parameters[_rfm_m]` returns (or an unpickled one), upload its level-1 draws once and reduce them on the GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .sampler import Sampler


class _DrawsOnDevice:
    """A bare handle (no data, no state) holding uploaded draws."""

    def __init__(self, draws, device=0):
        lvl1 = draws["level_1"]
        chains = len(lvl1)
        n_draws, N, ncol = lvl1[0].shape
        self.lib = L.load()
        self.h = C.c_void_p()
        self.N, self.chains, self.ncol, self.n_draws = N, chains, ncol, n_draws
        cfg = L.Config(model_dim=2 if ncol == 4 else 3, n_cov=1, n_chains=chains, chain_offset=0, n_mh_steps=0, rng_mode=L.RNG_FAST,
                       compat=0, sweep_mode=L.SWEEP_STREAM, device=int(device), reserved=0, n_local=N, n_global=N, gid_offset=0, seed=0)
        L.check(self.lib.clv_create(C.byref(self.h), C.byref(cfg)))
        try:
            a = np.ascontiguousarray(np.stack([np.asarray(c, dtype=np.float64) for c in lvl1]))
            L.check(self.lib.clv_upload_draws(self.h, L.dptr(a), n_draws), self.h)
        except Exception:
            self.close()
            raise

    def close(self):
        if self.h:
            self.lib.clv_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def posterior_summary(draws, mu_cap=0.05, device=0):
    """Per-customer posterior means and 2.5 / 97.5 % quantiles over all chains and draws -- the inputs of
    `compute_table4` / `post_mean_lambdas` / `post_mean_mus` (analysis_bi_helpers.py:15-27, 75-110).
    Returns a dict of (N,) arrays keyed by Sampler.SUMMARY_COLUMNS."""
    with _DrawsOnDevice(draws, device) as d:
        out = np.empty((d.N, len(Sampler.SUMMARY_COLUMNS)))
        L.check(d.lib.clv_posterior_summary(d.h, float(mu_cap), L.dptr(out)), d.h)
    return {k: out[:, j].copy() for j, k in enumerate(Sampler.SUMMARY_COLUMNS)}


def weekly_tracking(draws, birth_week, times, seed=0, device=0):
    """Posterior-predictive weekly incremental repeat transactions, averaged over all draws (Figure 2,
    bivariate/analysis_abe.py:446-464); `np.cumsum` of the result is the HB tracking curve."""
    b = np.ascontiguousarray(birth_week, dtype=np.float64)
    t = np.ascontiguousarray(times, dtype=np.float64)
    with _DrawsOnDevice(draws, device) as d:
        if b.size != d.N:
            raise ValueError("birth_week and the draws disagree on the number of customers")
        out = np.empty(t.size)
        L.check(d.lib.clv_weekly_tracking(d.h, L.dptr(b), L.dptr(t), int(t.size), int(seed) & 0xFFFFFFFFFFFFFFFF, L.dptr(out)), d.h)
    return out


def table4_inputs(draws, t_star=39.0, mu_cap=0.05, device=0):
    """The per-customer columns of the reference's Table 4 (analysis_bi_helpers.py:75-140), computed from the device
    summary with the reference's formulas: expected lifetime 1/mu/52, 1-year survival exp(-52 mu),
    E[x*] = P(alive) * lambda/mu * (1 - exp(-mu t*))."""
    s = posterior_summary(draws, mu_cap, device)
    lam, mu, z = s["mean_lambda"], s["mean_mu_capped"], s["p_alive"]
    with np.errstate(divide="ignore"):
        life = np.where(mu > 0, (1.0 / mu) / 52.0, np.inf)
    return {"Mean(λ)": lam, "2.5% tile λ": s["lambda_2.5"], "97.5% tile λ": s["lambda_97.5"], "Mean(μ)": mu,
            "2.5% tile μ": s["mu_2.5"], "97.5% tile μ": s["mu_97.5"], "Mean exp lifetime (yrs)": life,
            "Survival rate (1yr)": np.exp(-mu * 52), "P(alive at T_cal)": z,
            "Exp # of trans in val period": z * (lam / mu) * (1.0 - np.exp(-mu * t_star))}
