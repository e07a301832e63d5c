"""Pin the oracle (oracle/abe_oracle.py) to outputs of the reference itself
(tests/golden/*.npz, produced by tests/golden/make_golden.py from /root/reference)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import abe_oracle as ao
from oracle.streams import NumpyOrderStreams, ReplayStreams


def _cbs(d, cov=(), with_log_s=False):
    # F-order on purpose: the reference's `cbs[cols].to_numpy(float)` (bi:470) is an F-ordered view, BLAS
    # sums in a layout-dependent order, and the sampler amplifies 1-ulp differences ~1.8x per sweep while
    # Sigma_00 is still degenerate -- bit-equal inputs are needed for 1e-9 agreement after 60 sweeps.
    X = np.asfortranarray(np.column_stack([np.ones(len(d["x"]))] + [d[c].astype(float) for c in cov]))
    return ao.Cbs(x=d["x"].astype(np.int64), t_x=d["t_x"].astype(float), T_cal=d["T_cal"].astype(float),
                  X=X, log_s=d["log_s"].astype(float) if with_log_s else None)


def test_kat_bivariate_m1_real_rng(cdnow_abe):
    """Reference chain with its real PCG64 stream, seed 42."""
    g = load_golden("kat_bi_m1.npz")
    cbs = _cbs(cdnow_abe)
    out = ao.run_chain(cbs, ao.default_hyper(1, 2), NumpyOrderStreams(np.random.default_rng(42)),
                       mcmc=100, burnin=100, thin=1, D=2)
    np.testing.assert_allclose(out["level_2"][-1],
                               [-3.42062716, -3.21472713, 0.94580545, -0.26999558, 1.02892563], rtol=0, atol=1e-8)
    np.testing.assert_allclose(out["level_2"], g["level_2"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(out["level_1"][-1], g["level_1_last"], rtol=1e-10, atol=1e-12)
    assert abs(np.mean(out["log_likelihood"]) - (-6.05127448831035)) < 1e-10


def test_kat_bivariate_m2_two_chains(cdnow_abe):
    g = load_golden("kat_bi_m2.npz")
    cbs = _cbs(cdnow_abe, ["first_sales_scaled"])
    ll = []
    for ch in range(2):
        out = ao.run_chain(cbs, ao.default_hyper(2, 2), NumpyOrderStreams(np.random.default_rng(7 + ch)),
                           mcmc=40, burnin=20, thin=2, D=2)
        np.testing.assert_allclose(out["level_2"], g["level_2"][ch], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(out["level_1"][-1], g["level_1_last"][ch], rtol=1e-9, atol=1e-11)
        ll.append(out["log_likelihood"])
    assert abs(np.mean(np.concatenate(ll)) - float(g["loglik"])) < 1e-10


def test_kat_trivariate(cdnow_abe):
    g = load_golden("kat_tri.npz")
    cbs = _cbs(cdnow_abe, ["gender_F", "age_scaled"], with_log_s=True)
    out = ao.run_chain(cbs, ao.default_hyper(3, 3), NumpyOrderStreams(np.random.default_rng(11)),
                       mcmc=30, burnin=20, thin=1, D=3)
    np.testing.assert_allclose(out["level_2"], g["level_2"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(out["level_1"][-1], g["level_1_last"], rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("name", ["inj_bi_k1", "inj_bi_k2", "inj_bi_k4", "inj_tri_k3", "inj_tri_k1", "inj_edge"])
def test_injected_trajectories(name):
    """Reference `_run_chain` replayed with injected streams == oracle with the same streams."""
    g = load_golden(name + ".npz")
    D, S = int(g["D"]), int(g["S"])
    cbs = ao.Cbs(x=g["x"], t_x=g["t_x"], T_cal=g["T_cal"], X=g["X"], log_s=g.get("log_s"))
    T = g["u_z"].shape[0]
    out = ao.run_chain(cbs, ao.default_hyper(cbs.K, D), ReplayStreams(g), mcmc=T, burnin=0, thin=1, D=D,
                       n_mh_steps=S)
    np.testing.assert_array_equal(out["level_1"][:, :, 3], g["level_1"][:, :, 3])        # z bit-exact
    np.testing.assert_allclose(out["level_1"], g["level_1"], rtol=1e-9, atol=1e-300)
    np.testing.assert_allclose(out["level_2"], g["level_2"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(out["log_likelihood"], g["loglik"], rtol=1e-9)
    # scipy's chi-square degrees of freedom, as observed by the stub
    nu_n = ao.default_hyper(cbs.K, D)["nu_00"] + cbs.N
    np.testing.assert_array_equal(g["chi2_df_first"], [nu_n - D + 1 + i for i in range(D)])


def test_forecast_bivariate():
    g = load_golden("fc_bi.npz")
    xs = ao.forecast(g["T_cal"], g["level_1"], float(g["T_star"]), g["u"])
    np.testing.assert_array_equal(xs, g["x_star"])


def test_forecast_trivariate_spend():
    g = load_golden("fc_tri.npz")
    xs, sp = ao.forecast(g["T_cal"], g["level_1"], float(g["T_star"]), g["u"], eps=list(g["eps"]),
                         sigma_s=float(g["sigma_s"]))
    np.testing.assert_array_equal(xs, g["x_star"])
    np.testing.assert_allclose(sp, g["spend"], rtol=1e-12)


def test_poisson_inversion_matches_distribution():
    rng = np.random.default_rng(0)
    for m in (0.0, 0.3, 4.0, 37.5, 250.0):
        x = ao.poisson_inversion(np.full(200000, m), rng.random(200000))
        assert abs(x.mean() - m) < 5 * np.sqrt(max(m, 1e-9) / 200000) + 1e-12
        assert abs(x.var() - m) < 0.05 * max(m, 1e-9) + 1e-12


def test_philox_contract_variates_follow_the_reference_laws():
    """The device RNG contract restated in oracle/philox_np.py draws the Metropolis proposals from ONE Philox block per
    step: Student-t(3) from two uniforms (t / sqrt(3) = cos th / sqrt(sin^2 th + V / (1 - V)), the ratio of the two
    exponentials of the textbook construction being V / (1 - V)), accept uniform from the low bytes of the same words.
    Check the laws the reference draws from (`rng.standard_t(3)`, `rng.random()`, bi:316-330) and the independence the
    word sharing must not break."""
    from scipy import stats
    from oracle import philox_np as px
    n, S = 400_000, 4
    v = px.sampler_variates(20240229, 3, np.arange(n), 17, S, with_eta=True)
    ref_t = np.random.default_rng(0).standard_t(3, n)
    for s in range(S):
        for k in ("t3_l", "t3_m"):
            assert stats.kstest(v[k][s], stats.t(3).cdf).pvalue > 1e-3, (k, s)
            assert stats.ks_2samp(v[k][s], ref_t).pvalue > 1e-3, (k, s)
            # heavy tails are there (a 24-bit V reaches |t| in the thousands)
            assert abs(np.mean(np.abs(v[k][s]) > 10) / (2 * stats.t.sf(10, 3)) - 1) < 0.15
        assert stats.kstest(v["u_acc"][s], "uniform").pvalue > 1e-3
        # proposals for log lambda and log mu, and the accept uniform, are mutually independent
        assert abs(np.corrcoef(v["t3_l"][s], v["t3_m"][s])[0, 1]) < 0.01
        assert abs(np.corrcoef(np.abs(v["t3_l"][s]) < 1, np.abs(v["t3_m"][s]) < 1)[0, 1]) < 0.01
        for k in ("t3_l", "t3_m"):
            assert abs(np.corrcoef(v["u_acc"][s], np.abs(v[k][s]) < 1)[0, 1]) < 0.01
            assert abs(np.corrcoef(v["u_acc"][s], v[k][s] > 0)[0, 1]) < 0.01
    # consecutive steps and sweeps use different counters
    assert abs(np.corrcoef(v["t3_l"][0], v["t3_l"][1])[0, 1]) < 0.01
    w = px.sampler_variates(20240229, 3, np.arange(n), 18, S)
    assert abs(np.corrcoef(v["t3_l"][0], w["t3_l"][0])[0, 1]) < 0.01
    assert stats.kstest(v["n_eta"], "norm").pvalue > 1e-3 and stats.kstest(v["u_z"], "uniform").pvalue > 1e-3
