#!/usr/bin/env python
"""Reference posterior summaries for the statistical-parity tests (build container only; needs /root/reference).

Runs the UNMODIFIED reference sampler, one process per chain (chain c == chains=1, seed=seed+c, bi:486), and stores
pooled means, sds and Geyer MCSEs of every level-2 column plus a few level-1 aggregates:
    python tests/golden/make_posterior_golden.py tri_k3 | bi_k4
"""
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np
import pandas as pd

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

CASES = {
    "tri_k3": dict(model="tri", cov=["gender_F", "age_scaled"], chains=8, burnin=3000, mcmc=3000, seed=42),
    "bi_k4": dict(model="bi", cov=["first_sales_scaled", "age_scaled", "gender_binary"], chains=8, burnin=4000, mcmc=4000, seed=42),
    # BASELINE.json configs[1] / configs[2]: full CDNOW (23 570 customers).  Burn-in covers the degenerate start.
    # The reference drivers' real settings (run_mcmc_full.py:137-147, trivariate/run_mcmc_full.py:80-90): 10 000 + 4 000, thin 1.
    "c2_full_bi_k2": dict(model="bi", data="cdnow_full.npz", cov=["first_sales_scaled"], chains=4, burnin=10000, mcmc=4000, seed=42),
    "c3_full_tri_k3": dict(model="tri", data="cdnow_full.npz", cov=["gender_F", "age_scaled"], chains=4, burnin=10000, mcmc=4000, seed=42),
}


def load(name="cdnow_abe.npz"):
    d = np.load(os.path.join(HERE, name))
    return pd.DataFrame({k: d[k] for k in d.files})


def one_chain(args):
    name, c = args
    cfg = CASES[name]
    sys.path.insert(0, REF)
    cbs = load(cfg.get("data", "cdnow_abe.npz"))
    if cfg["model"] == "tri":
        import src.models.trivariate.mcmc as m
        out = m.mcmc_draw_parameters_rfm_m(cbs, covariates=cfg["cov"], mcmc=cfg["mcmc"], burnin=cfg["burnin"], thin=1, chains=1,
                                           seed=cfg["seed"] + c, trace=0)
    else:
        import src.models.bivariate.mcmc as m
        out = m.mcmc_draw_parameters(cbs, covariates=cfg["cov"], mcmc=cfg["mcmc"], burnin=cfg["burnin"], thin=1, chains=1,
                                     seed=cfg["seed"] + c, trace=0)
    l1 = out["level_1"][0]
    # level-1 aggregates: column means over every 20th draw, and per-customer posterior means of the first 8 customers
    return out["level_2"][0], l1[::20].mean(axis=(0, 1)), out["log_likelihood"], l1[:, :8, :].mean(axis=0)


def main():
    name = sys.argv[1]
    cfg = CASES[name]
    with ProcessPoolExecutor(min(8, cfg["chains"])) as ex:
        res = list(ex.map(one_chain, [(name, c) for c in range(cfg["chains"])]))
    sys.path.insert(0, ROOT)
    from mcmc_clv_model_b200.diagnostics import summarize
    l2 = np.stack([r[0] for r in res])
    s = summarize(l2)
    np.savez_compressed(os.path.join(HERE, f"post_{name}.npz"),
                        mean=np.array([s[j]["mean"] for j in range(l2.shape[2])]),
                        sd=np.array([s[j]["sd"] for j in range(l2.shape[2])]),
                        mcse=np.array([s[j]["mcse_mean"] for j in range(l2.shape[2])]),
                        ess=np.array([s[j]["ess_geyer"] for j in range(l2.shape[2])]),
                        chain_means=l2.mean(axis=1), level1_col_means=np.mean([r[1] for r in res], axis=0),
                        quantiles=np.percentile(l2.reshape(-1, l2.shape[2]), [2.5, 50, 97.5], axis=0),
                        level_2=l2.astype(np.float32), customer_means8=np.mean([r[3] for r in res], axis=0),
                        loglik=np.mean([r[2] for r in res]), chains=cfg["chains"], burnin=cfg["burnin"], mcmc=cfg["mcmc"],
                        covariates=np.array(cfg["cov"]))
    print(name, "mean", np.round([s[j]["mean"] for j in range(l2.shape[2])], 3))
    print(name, "mcse", np.round([s[j]["mcse_mean"] for j in range(l2.shape[2])], 4))


if __name__ == "__main__":
    main()
