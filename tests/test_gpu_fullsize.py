"""BASELINE.json's full sizes, checked through size-independent properties (the oracle cannot run 10 M customers in
seconds): C4 = 10 M synthetic customers x 4 covariates; C5-shaped forecast on resident draws."""
import numpy as np
import pytest

from mcmc_clv_model_b200 import Sampler
from mcmc_clv_model_b200.synthetic import C4_BETA, C4_GAMMA, C4_SEED, C4_T_CAL, generate_cbs_arrays

pytestmark = [pytest.mark.gpu, pytest.mark.slow]


def test_c4_full_size_properties():
    n = 10_000_000
    c = generate_cbs_arrays(n, C4_BETA, C4_GAMMA, T_cal=C4_T_CAL, T_star=39.0, seed=C4_SEED, with_truth=True)
    assert np.all((c["x"] == 0) == (c["t_x"] == 0)) and np.all(c["t_x"] <= c["T_cal"])
    outs = []
    for rep in range(2):
        with Sampler(c["x"], c["t_x"], c["T_cal"], c["X"], chains=1, seed=4) as s:
            s.set_state(0, log_lambda=np.log(c["lambda_true"]), log_mu=np.log(c["mu_true"]), beta=C4_BETA, Sigma=C4_GAMMA)
            out = s.run(3, 2, 1)
            st = s.get_state(0)
            stats = s.init_stats
        outs.append(out)
    # determinism: the same seed gives the same 10 M-customer draws, bit for bit (atomics on int64 fixed point)
    np.testing.assert_array_equal(outs[0]["level_1"], outs[1]["level_1"])
    np.testing.assert_array_equal(outs[0]["level_2"], outs[1]["level_2"])
    l1 = out["level_1"][0][-1]
    lam, mu, tau, z = l1.T
    # support of every block's output
    assert set(np.unique(z)) <= {0.0, 1.0}
    assert np.all(tau[z == 1] > c["T_cal"][z == 1])                                  # bi:217: alive => tau > T_cal
    ch = z == 0
    assert np.all((tau[ch] >= c["t_x"][ch] - 1e-9) & (tau[ch] <= c["T_cal"][ch] + 1e-9))   # bi:219-226: churned in [t_x, T_cal]
    assert np.all((lam > 0) & (mu > 0)) and np.all(np.abs(np.log(lam)) <= 70) and np.all(np.log(mu) <= 5.0 + 1e-12)
    # the kept draw is the state: lambda == exp(log_lambda) of get_state, z/tau equal the stored ones
    np.testing.assert_allclose(lam, np.exp(st["log_lambda"]), rtol=1e-14)
    np.testing.assert_array_equal(z, st["z"])
    np.testing.assert_array_equal(tau, st["tau"])
    # per-draw log-likelihood (bi:423-428) recomputed from the draw itself
    lik = c["x"] * np.log(lam) + (1 - z) * np.log(mu) - (lam + mu) * (z * c["T_cal"] + (1 - z) * tau)
    assert abs(out["loglik_sum"][0, -1] / lik.sum() - 1) < 1e-9
    # level-2 draw of the last sweep is consistent with the state it was drawn from: with N = 1e7 the posterior of
    # (beta, Sigma) is within ~1e-3 of the regression of the log-parameters on X
    Y = np.column_stack([st["log_lambda"], st["log_mu"]])
    B = np.linalg.solve(c["X"].T @ c["X"], c["X"].T @ Y)
    np.testing.assert_allclose(st["beta"], B, atol=0.02)
    np.testing.assert_allclose(st["Sigma"], np.cov((Y - c["X"] @ B).T), rtol=0.02, atol=0.01)
    # exact initialisation statistics at this size
    lam0 = c["x"].mean() / np.mean(np.where(c["t_x"] == 0, c["T_cal"], c["t_x"]))
    assert abs(stats["lam_init"] / lam0 - 1) < 1e-12
    # chains started at the generating parameters stay there: Sigma_00 ~ Gamma_00
    assert abs(st["Sigma"][0, 0] - C4_GAMMA[0, 0]) < 0.1 and abs(st["beta"][0, 0] - C4_BETA[0, 0]) < 0.1


def test_c5_forecast_resident_properties():
    n, nd = 1_000_000, 100
    c = generate_cbs_arrays(n, C4_BETA, C4_GAMMA, T_cal=C4_T_CAL, T_star=39.0, seed=C4_SEED + 1, with_truth=True)
    with Sampler(c["x"], c["t_x"], c["T_cal"], c["X"], chains=1, seed=7) as s:
        s.set_state(0, log_lambda=np.log(c["lambda_true"]), log_mu=np.log(c["mu_true"]), beta=C4_BETA, Sigma=C4_GAMMA)
        s.run_resident(10, nd, 1)
        a = s.forecast_resident(T_star=39.0, seed=42)
        b = s.forecast_resident(T_star=39.0, seed=42)
        z = s.forecast_resident(T_star=0.0, seed=42)
        summ_alive = s.posterior_summary()["p_alive"]
    np.testing.assert_array_equal(a["mean_x_star"], b["mean_x_star"])               # deterministic
    assert np.all(z["mean_x_star"] == 0)                                             # zero horizon => no transactions
    np.testing.assert_allclose(a["p_alive"], summ_alive, rtol=1e-12)                 # two kernels, one quantity
    assert np.all(a["mean_x_star"][a["p_alive"] == 0] <= 39.0 * 50)                  # churned: bounded by the clipped horizon
    # aggregate calibration of the simulation against the generated hold-out purchases
    assert abs(a["mean_x_star"].mean() / c["x_star"].mean() - 1) < 0.35
