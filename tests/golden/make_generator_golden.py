#!/usr/bin/env python
"""Golden moments of the reference's OWN `generate_pareto_abe` (bivariate/mcmc.py:95-187), run unmodified in the build
container (needs /root/reference):  python tests/golden/make_generator_golden.py
Two cases, n = 20 000, seed 42: scalar T_cal, and a vector T_cal (which also pins the reference's CBS convention for
cohorts: t_x on the shifted clock, one scalar T_cal = max(T_cal), bi:165)."""
import os
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
BETA = np.array([[-3.5, -3.6], [0.2, -0.1]])
GAMMA = np.array([[1.4, 0.3], [0.3, 2.5]])
N, T_STAR = 20000, 39.0


def moments(cbs, T_rel, t_rel):
    """Statistics of one CBS (relative clock): value and standard error."""
    x, xs = cbs["x"].to_numpy(float), cbs["x_star"].to_numpy(float)
    alive = cbs["alive_true"].to_numpy(float)
    stats = {
        "p_x0": x == 0, "p_x_ge1": x >= 1, "p_x_ge3": x >= 3, "p_x_ge10": x >= 10,
        "mean_sqrt_x": np.sqrt(x), "mean_log1p_x": np.log1p(x),
        "p_xs0": xs == 0, "p_xs_ge3": xs >= 3, "mean_log1p_xs": np.log1p(xs),
        "alive": alive, "mean_tx_over_T": t_rel / T_rel, "p_tx_late": (t_rel / T_rel) > 0.75,
        "mean_log_lambda": np.log(cbs["lambda_true"].to_numpy()), "mean_log_mu": np.log(cbs["mu_true"].to_numpy()),
        "corr_proxy": np.log1p(x) * alive,
    }
    names = sorted(stats)
    val = np.array([np.mean(stats[k].astype(float)) for k in names])
    se = np.array([np.std(stats[k].astype(float), ddof=1) / np.sqrt(len(x)) for k in names])
    return names, val, se


def main():
    sys.path.insert(0, REF)
    import src.models.bivariate.mcmc as m
    out = {}
    # scalar T_cal
    cbs, elog = m.generate_pareto_abe(N, 32.0, T_STAR, BETA, GAMMA, seed=42)
    names, val, se = moments(cbs, np.full(N, 32.0), cbs["t_x"].to_numpy())
    out.update(names=np.array(names), scalar_val=val, scalar_se=se, scalar_T_cal=cbs["T_cal"].to_numpy()[:8],
               scalar_x_hist=np.bincount(np.minimum(cbs["x"].to_numpy(), 30), minlength=31))
    # vector T_cal, covariates given
    rng = np.random.default_rng(7)
    T_vec = rng.uniform(27.0, 38.857142857142854, N)
    cov = rng.uniform(-1, 1, N)
    cbs2, elog2 = m.generate_pareto_abe(N, T_vec, T_STAR, BETA, GAMMA, covars=cov, seed=43)
    T_fix = T_vec.max()
    T_zero = T_fix - T_vec
    t_rel = cbs2["t_x"].to_numpy() - T_zero                      # the reference's t_x is on the shifted clock (bi:158,165)
    assert np.all(cbs2["T_cal"].to_numpy() == T_fix) and np.all(t_rel > -1e-9)
    names2, val2, se2 = moments(cbs2, T_vec, np.maximum(t_rel, 0.0))
    assert names2 == names
    out.update(vector_val=val2, vector_se=se2, vector_T_cal_in=T_vec, vector_cov=cov, vector_T_fix=T_fix,
               vector_min_tx_minus_Tzero=float(t_rel.min()), beta=BETA, gamma=GAMMA, n=N, T_star=T_STAR,
               elog_first_purchase_is_Tzero=bool(np.allclose(elog2.groupby("cust")["t"].min().to_numpy(), T_zero)))
    np.savez_compressed(os.path.join(HERE, "gen_ref.npz"), **out)
    for k, a, b in zip(names, val, val2):
        print(f"{k:18s} scalar {a:9.5f}   vector {b:9.5f}")


if __name__ == "__main__":
    main()
