"""Known-answer tests of the convergence diagnostics that replace az.summary / az.plot_autocorr
(bivariate/analysis_abe.py:651-706) -- the ESS/s metric of bench.py stands on them."""
import numpy as np
import pytest

from mcmc_clv_model_b200 import diagnostics as dg


def _ar1(rng, chains, n, rho, shift=None):
    e = rng.normal(size=(chains, n)) * np.sqrt(1.0 - rho * rho)
    x = np.empty((chains, n))
    x[:, 0] = rng.normal(size=chains)
    for t in range(1, n):
        x[:, t] = rho * x[:, t - 1] + e[:, t]
    if shift is not None:
        x += np.asarray(shift)[:, None]
    return x


def test_iid_draws_have_ess_about_n_and_rhat_one():
    rng = np.random.default_rng(1)
    x = rng.normal(size=(4, 4000))
    n = x.size
    assert 0.85 * n < dg.ess_bulk(x) < 1.15 * n
    assert 0.80 * n < dg.ess_tail(x) < 1.20 * n
    assert 0.85 * n < dg.ess_geyer(x) < 1.15 * n
    assert 0.85 * n < dg.ess_mean(x) < 1.15 * n
    assert dg.rhat(x) < 1.01


@pytest.mark.parametrize("rho", [0.5, 0.9, 0.97])
def test_ar1_ess_matches_the_closed_form(rho):
    """AR(1): ESS = n (1 - rho) / (1 + rho)."""
    rng = np.random.default_rng(int(rho * 100))
    x = _ar1(rng, 4, 20000, rho)
    expect = x.size * (1 - rho) / (1 + rho)
    for f in (dg.ess_bulk, dg.ess_geyer, dg.ess_mean):
        assert 0.75 * expect < f(x) < 1.30 * expect, (f.__name__, f(x), expect)
    assert dg.ess_tail(x) < x.size                       # tails of a positively correlated chain mix no better than iid
    assert dg.rhat(x) < 1.02


def test_autocorrelation_function_of_ar1():
    rng = np.random.default_rng(3)
    rho = 0.8
    ac = dg.autocorr(_ar1(rng, 2, 50000, rho), max_lag=10)
    assert ac.shape == (2, 11)
    np.testing.assert_allclose(ac[:, 0], 1.0)
    np.testing.assert_allclose(ac, np.broadcast_to(rho ** np.arange(11), ac.shape), atol=0.03)


def test_rhat_flags_chains_that_have_not_mixed():
    rng = np.random.default_rng(4)
    assert dg.rhat(_ar1(rng, 4, 2000, 0.3, shift=[0, 0, 1.5, 1.5])) > 1.1
    # same location, different scale: only the folded (tail) R-hat sees it
    y = rng.normal(size=(4, 2000)) * np.array([1, 1, 4, 4])[:, None]
    assert dg.rhat(y) > 1.1
    # shifted chains also collapse the multi-chain bulk ESS
    assert dg.ess_bulk(_ar1(rng, 4, 2000, 0.3, shift=[0, 0, 3, 3])) < 50


def test_bulk_ess_and_rhat_are_invariant_under_monotone_transforms():
    rng = np.random.default_rng(5)
    x = _ar1(rng, 3, 3000, 0.7)
    assert dg.ess_bulk(x) == pytest.approx(dg.ess_bulk(np.exp(3 * x)), rel=1e-12)
    assert dg.ess_tail(x) == pytest.approx(dg.ess_tail(np.exp(3 * x)), rel=1e-12)


def test_summary_table_has_the_az_summary_columns():
    rng = np.random.default_rng(6)
    l2 = np.stack([_ar1(rng, 2, 1500, r).T for r in (0.2, 0.6, 0.9)], axis=0).transpose(2, 1, 0)   # (chains, draws, P)
    assert l2.shape == (2, 1500, 3)
    t = dg.summary_table(l2, names=["a", "b", "c"])
    assert list(t.columns) == ["mean", "sd", "hdi_3%", "hdi_97%", "mcse_mean", "ess_bulk", "ess_tail", "r_hat"]
    assert list(t.index) == ["a", "b", "c"]
    assert t["ess_bulk"]["a"] > t["ess_bulk"]["b"] > t["ess_bulk"]["c"]
    lo, hi = dg.hdi(rng.normal(size=200000))
    assert lo == pytest.approx(-1.88, abs=0.03) and hi == pytest.approx(1.88, abs=0.03)      # 94 % of a standard normal
    assert dg.min_ess(l2, "bulk") == pytest.approx(t["ess_bulk"].min(), rel=1e-3)
