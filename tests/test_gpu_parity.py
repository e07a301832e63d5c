"""GPU parity tests proper: the CUDA path (through the C-ABI) against the reference's own outputs
(tests/golden/*.npz) and against the oracle on seeded inputs.

Tolerances (BASELINE.json north_star): driven by identical injected streams from the same state,
z and x* bit-exact, continuous updates 1e-6 relative."""
import numpy as np
import pytest

from conftest import load_golden
from mcmc_clv_model_b200 import Sampler
from oracle import abe_oracle as ao
from oracle.streams import PhiloxStreams, ReplayStreams

pytestmark = pytest.mark.gpu

RTOL = 1e-6   # continuous updates, injected streams (north_star)


def _sweep_arrays(g, t, D):
    names = ["u_z", "e_tau", "u_tau", "t3_l", "t3_m", "u_acc", "iw_norm", "iw_chi2", "beta_norm"] + (["n_eta"] if D == 3 else [])
    return {n: g[n][t][None] for n in names}


@pytest.mark.parametrize("name", ["inj_bi_k1", "inj_bi_k2", "inj_bi_k4", "inj_tri_k3", "inj_tri_k1", "inj_edge"])
def test_injected_trajectory_vs_reference(name):
    """Reference `_run_chain` outputs (golden) vs the CUDA sweep fed the same injected streams."""
    g = load_golden(name + ".npz")
    D, S = int(g["D"]), int(g["S"])
    T = g["u_z"].shape[0]
    with Sampler(g["x"], g["t_x"], g["T_cal"], g["X"], g.get("log_s"), model_dim=D, chains=1, n_mh_steps=S,
                 rng="injected") as s:
        for t in range(T):
            out = s.sweep_injected(_sweep_arrays(g, t, D), keep=True)
            l1, ref = out["level_1"][0], g["level_1"][t]
            np.testing.assert_array_equal(l1[:, 3], ref[:, 3], err_msg=f"z differs at sweep {t}")
            np.testing.assert_allclose(l1, ref, rtol=RTOL, atol=0, err_msg=f"level_1 sweep {t}")
            np.testing.assert_allclose(out["level_2"][0], g["level_2"][t], rtol=RTOL, atol=1e-9, err_msg=f"level_2 sweep {t}")
            np.testing.assert_allclose(out["loglik_sum"][0] / len(g["x"]), g["loglik"][t], rtol=RTOL)


def test_injected_two_chains_match_single_chain_runs():
    """Chains are independent: a 2-chain handle equals two 1-chain handles."""
    g = load_golden("inj_bi_k2.npz")
    rng = np.random.default_rng(5)
    N, S, K, D = len(g["x"]), 20, 2, 2
    from oracle.streams import random_replay_arrays
    a = [random_replay_arrays(rng, 2, N, S, D, K, 5 + N) for _ in range(2)]
    both = {k: np.stack([a[0][k], a[1][k]], axis=1) for k in a[0]}     # (T, chains, ...)
    with Sampler(g["x"], g["t_x"], g["T_cal"], g["X"], model_dim=2, chains=2, n_mh_steps=S, rng="injected") as s2:
        outs2 = [s2.sweep_injected({k: v[t] for k, v in both.items()}) for t in range(2)]
    for c in range(2):
        with Sampler(g["x"], g["t_x"], g["T_cal"], g["X"], model_dim=2, chains=1, n_mh_steps=S, rng="injected") as s1:
            for t in range(2):
                o = s1.sweep_injected({k: v[t][None] for k, v in a[c].items()})
                np.testing.assert_array_equal(o["level_1"][0], outs2[t]["level_1"][c])
                np.testing.assert_array_equal(o["level_2"][0], outs2[t]["level_2"][c])


@pytest.mark.parametrize("D,cov", [(2, []), (2, ["first_sales_scaled"]), (3, ["gender_F", "age_scaled"])])
def test_strict_philox_trajectory_vs_oracle(cdnow_abe, D, cov):
    """The production sweep loop (clv_run, strict-f64 Philox) against the oracle replaying the same
    counter-based variates: the whole device pipeline (RNG transforms, fixed-point level-2 statistics,
    Bartlett/IW draw, burn-in/thin bookkeeping) over several sweeps."""
    d = cdnow_abe
    n = 700
    X = np.column_stack([np.ones(n)] + [d[c][:n].astype(float) for c in cov])
    cbs = ao.Cbs(x=d["x"][:n].astype(np.int64), t_x=d["t_x"][:n], T_cal=d["T_cal"][:n], X=X,
                 log_s=d["log_s"][:n] if D == 3 else None)
    seed, burnin, mcmc, thin, S = 1234, 2, 5, 2, 20
    for chain in (0, 1):
        ora = ao.run_chain(cbs, ao.default_hyper(cbs.K, D), PhiloxStreams(seed, chain, np.arange(n), S, D, cbs.K),
                           mcmc=mcmc, burnin=burnin, thin=thin, D=D, n_mh_steps=S)
        with Sampler(cbs.x, cbs.t_x, cbs.T_cal, X, cbs.log_s, model_dim=D, chains=1, chain_offset=chain,
                     n_mh_steps=S, seed=seed, rng="strict") as s:
            out = s.run(burnin, mcmc, thin)
        np.testing.assert_array_equal(out["level_1"][0][:, :, 3], ora["level_1"][:, :, 3])
        np.testing.assert_allclose(out["level_1"][0], ora["level_1"], rtol=RTOL)
        np.testing.assert_allclose(out["level_2"][0], ora["level_2"], rtol=RTOL, atol=1e-9)
        np.testing.assert_allclose(out["loglik_sum"][0] / n, ora["log_likelihood"], rtol=RTOL)


@pytest.mark.parametrize("S", [1, 2, 5, 7])
def test_strict_philox_odd_step_counts_vs_oracle(cdnow_abe, S):
    """n_mh_steps other than the default (the eta slot 1 + 2S moves with it).  Both sweep paths against the oracle
    replay."""
    d = cdnow_abe
    n = 300
    X = np.column_stack([np.ones(n), d["first_sales_scaled"][:n]])
    cbs = ao.Cbs(x=d["x"][:n].astype(np.int64), t_x=d["t_x"][:n], T_cal=d["T_cal"][:n], X=X, log_s=d["log_s"][:n])
    for D in (2, 3):
        ora = ao.run_chain(cbs, ao.default_hyper(cbs.K, D), PhiloxStreams(77, 0, np.arange(n), S, D, cbs.K),
                           mcmc=4, burnin=3, thin=1, D=D, n_mh_steps=S)
        for mode in ("stream", "persistent"):
            with Sampler(cbs.x, cbs.t_x, cbs.T_cal, X, cbs.log_s if D == 3 else None, model_dim=D, n_mh_steps=S, seed=77,
                         rng="strict", sweep_mode=mode) as s:
                out = s.run(3, 4, 1)
            np.testing.assert_array_equal(out["level_1"][0][:, :, 3], ora["level_1"][:, :, 3])
            np.testing.assert_allclose(out["level_1"][0], ora["level_1"], rtol=RTOL)
            np.testing.assert_allclose(out["level_2"][0], ora["level_2"], rtol=RTOL, atol=1e-9)


def test_chain_offset_reproduces_chain(cdnow_abe):
    """Any chain of a run can be reproduced alone (`chains=1, chain_offset=c`): the counterpart of the reference's
    `chains=1, seed=seed+ch` (bi:486) under a counter-based RNG whose chain index travels in the counter."""
    d = cdnow_abe
    n = 500
    X = np.ones((n, 1))
    args = (d["x"][:n], d["t_x"][:n], d["T_cal"][:n], X)
    with Sampler(*args, chains=3, seed=42) as s:
        a = s.run(3, 4, 1)
    with Sampler(*args, chains=3, seed=42) as s:
        b = s.run(3, 4, 1)          # and runs are deterministic
    with Sampler(*args, chains=1, chain_offset=2, seed=42) as s:
        c = s.run(3, 4, 1)
    np.testing.assert_array_equal(a["level_2"], b["level_2"])
    np.testing.assert_array_equal(a["level_1"], b["level_1"])
    np.testing.assert_array_equal(a["level_2"][2], c["level_2"][0])
    np.testing.assert_array_equal(a["level_1"][2], c["level_1"][0])
    assert not np.array_equal(a["level_2"][0], a["level_2"][1])


def test_forecast_bivariate_vs_reference():
    g = load_golden("fc_bi.npz")
    import ctypes as C
    from mcmc_clv_model_b200 import _lib as L
    lib = L.load()
    l1 = np.ascontiguousarray(g["level_1"])
    nd, N, nc = l1.shape
    xs = np.empty((nd, N), dtype=np.int64)
    cfg = L.ForecastConfig(device=0, ncol=nc, n_draws_total=nd, n_customers=N, T_star=float(g["T_star"]), seed=1)
    L.check(lib.clv_forecast_injected(C.byref(cfg), L.dptr(l1), L.dptr(np.ascontiguousarray(g["T_cal"])),
                                      L.dptr(np.ascontiguousarray(g["u"])), None, 0, None,
                                      xs.ctypes.data_as(L.c_int64_p), None))
    np.testing.assert_array_equal(xs, g["x_star"])          # x* bit-exact


def test_forecast_trivariate_spend_vs_reference():
    g = load_golden("fc_tri.npz")
    import ctypes as C
    from mcmc_clv_model_b200 import _lib as L
    lib = L.load()
    l1 = np.ascontiguousarray(g["level_1"])
    nd, N, nc = l1.shape
    stride = g["eps"].shape[1]
    # the reference consumes eps[d][:sum x*] in customer order: offsets = exclusive cumsum of x* within the draw
    off = (np.cumsum(g["x_star"], axis=1) - g["x_star"]) + (np.arange(nd) * stride)[:, None]
    off = np.ascontiguousarray(off, dtype=np.int64)
    eps = np.ascontiguousarray(g["eps"]).ravel()
    xs = np.empty((nd, N), dtype=np.int64)
    sp = np.empty((nd, N))
    cfg = L.ForecastConfig(device=0, ncol=nc, n_draws_total=nd, n_customers=N, T_star=float(g["T_star"]), seed=1,
                           simulate_spend=1, sigma_s=float(g["sigma_s"]))
    L.check(lib.clv_forecast_injected(C.byref(cfg), L.dptr(l1), L.dptr(np.ascontiguousarray(g["T_cal"])),
                                      L.dptr(np.ascontiguousarray(g["u"])), L.dptr(eps), eps.size,
                                      off.ctypes.data_as(L.c_int64_p), xs.ctypes.data_as(L.c_int64_p), L.dptr(sp)))
    np.testing.assert_array_equal(xs, g["x_star"])
    np.testing.assert_allclose(sp, g["spend"], rtol=RTOL)


def test_forecast_philox_vs_oracle():
    """Production forecast (Philox uniforms) == oracle Poisson inversion on the same counter-based uniforms."""
    from oracle import philox_np as px
    from mcmc_clv_model_b200.api import _forecast
    g = load_golden("fc_bi.npz")
    l1 = g["level_1"]
    nd, N, _ = l1.shape
    l1 = l1.copy()
    l1[:, 5:9, 0] = [2.0, 7.5, 30.0, 400.0]          # means of 78 ... 15600 when alive: the PTRS branch
    l1[:, 5:9, 3] = 1.0
    x, _ = _forecast(g["T_cal"], [l1[:2], l1[2:]], 39.0, 77, False, 0.5)
    u = np.stack([px.forecast_uniform(77, np.arange(N), d) for d in range(nd)])
    ref = ao.forecast(g["T_cal"], l1, 39.0, u)
    m = l1[:, :, 0] * ao.future_horizon(g["T_cal"][None, :], l1[:, :, 2], l1[:, :, 3], 39.0)
    big = m >= px.PTRS_MIN_MEAN
    assert big.sum() >= 4 * nd
    for d, i in zip(*np.nonzero(big)):
        ref[d, i] = px.forecast_poisson_ptrs(77, int(i), int(d), float(m[d, i]))
    np.testing.assert_array_equal(x, ref)


def test_forecast_large_mean_distribution():
    """PTRS branch: mean and variance of x* for means 60 ... 5000."""
    from mcmc_clv_model_b200.api import _forecast
    N = 20000
    for m in (60.0, 93.7, 800.0, 5000.0):
        l1 = np.zeros((4, N, 4))
        l1[:, :, 0] = m / 39.0
        l1[:, :, 3] = 1.0
        x, _ = _forecast(np.full(N, 30.0), [l1], 39.0, 5, False, 0.5)
        n = x.size
        assert abs(x.mean() - m) < 5 * np.sqrt(m / n), (m, x.mean())
        assert abs(x.var() / m - 1) < 0.03, (m, x.var())


@pytest.mark.parametrize("D", [2, 3])
def test_device_init_statistics_equal_host_exact_sums(cdnow_full, monkeypatch, D):
    """clv_init_state(NULL): the device's exact integer sums == hostmath's, bit for bit."""
    from mcmc_clv_model_b200.hostmath import init_statistics
    d = cdnow_full
    X = np.column_stack([np.ones(d["x"].size), d["first_sales_scaled"], d["gender_F"].astype(float), d["age_scaled"]])
    log_s = d["log_s"] if D == 3 else None
    host = init_statistics(d["x"], d["t_x"], d["T_cal"], X, log_s, d["x"].size)
    for generic in ("0", "1"):        # K <= 5: sums kept in registers (k_init_quantities_t) / the general kernel
        monkeypatch.setenv("CLV_INIT_GENERIC", generic)
        with Sampler(d["x"], d["t_x"], d["T_cal"], X, log_s, model_dim=D, chains=1) as s:
            dev = s.init_stats
        for k in ("lam_init", "mean_mu_init", "mean_log_s", "omega2", "max_abs_x"):
            assert dev[k] == host[k], (k, dev[k], host[k], generic)
        np.testing.assert_array_equal(dev["xtx"], host["xtx"])
    # and close to the reference's plain NumPy means (bi:368-374)
    lam = d["x"].mean() / np.mean(np.where(d["t_x"] == 0, d["T_cal"], d["t_x"]))
    assert abs(dev["lam_init"] / lam - 1) < 1e-13


@pytest.mark.parametrize("D", [2, 3])
def test_persistent_kernel_equals_stream_mode(cdnow_abe, D):
    """The fused cooperative kernel (grid barrier per sweep) and the two-kernels-per-sweep path are the same chain,
    bit for bit, including draw bookkeeping across chunked level-1 storage and the trace segments."""
    import os
    d = cdnow_abe
    n = 2357
    X = np.column_stack([np.ones(n), d["first_sales_scaled"][:n]])
    log_s = d["log_s"][:n] if D == 3 else None
    outs = {}
    for mode in ("stream", "persistent"):
        with Sampler(d["x"][:n], d["t_x"][:n], d["T_cal"][:n], X, log_s, model_dim=D, chains=3, seed=5, sweep_mode=mode) as s:
            seen = []
            a = s.run(7, 23, 3, trace=5, progress=lambda st, tot: seen.append(st))
            assert seen == [5, 10, 15, 20, 25, 30]
            b = s.run(0, 4, 1)          # continues the same chains
            st = s.get_state(1)
        outs[mode] = (a, b, st)
    for k in ("level_1", "level_2", "loglik_sum"):
        np.testing.assert_array_equal(outs["stream"][0][k], outs["persistent"][0][k], err_msg=k)
        np.testing.assert_array_equal(outs["stream"][1][k], outs["persistent"][1][k], err_msg=k)
    for k in ("log_lambda", "log_mu", "z", "tau", "beta", "Sigma"):
        np.testing.assert_array_equal(outs["stream"][2][k], outs["persistent"][2][k], err_msg=k)
    # chunked draw storage (two device buffers, flushed while sweeps continue) gives the same arrays
    os.environ["CLV_DRAW_BUFFER_BYTES"] = str(2 * 3 * n * (4 if D == 2 else 5) * 8 * 2)   # 2 draws per chunk
    try:
        for mode in ("stream", "persistent"):
            with Sampler(d["x"][:n], d["t_x"][:n], d["T_cal"][:n], X, log_s, model_dim=D, chains=3, seed=5, sweep_mode=mode) as s:
                a = s.run(7, 23, 3)
            for k in ("level_1", "level_2", "loglik_sum"):
                np.testing.assert_array_equal(a[k], outs["stream"][0][k], err_msg=f"chunked {mode} {k}")
    finally:
        del os.environ["CLV_DRAW_BUFFER_BYTES"]


def test_staged_host_transfers_give_the_same_arrays():
    """Level-1 draws and forecast arrays travel through a ring of page-locked buffers with a parallel host memcpy when
    they are large (>= 64 MB); small ones take the driver's pageable path.  Both must deliver identical arrays, for
    every piece size / flush granularity, with one and with two device chunk buffers."""
    import os
    from mcmc_clv_model_b200.synthetic import C4_BETA, C4_GAMMA, C4_SEED, C4_T_CAL, generate_cbs_arrays
    from mcmc_clv_model_b200.api import _forecast
    n = 150_000
    c = generate_cbs_arrays(n, C4_BETA, C4_GAMMA, T_cal=C4_T_CAL, seed=C4_SEED, with_truth=False)
    def run(env):
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            with Sampler(c["x"], c["t_x"], c["T_cal"], c["X"], chains=2, seed=3, sweep_mode="stream") as s:
                out = s.run(2, 21, 2)                        # 11 draws x 2 chains x 4.8 MB
            xs, _ = _forecast(c["T_cal"], out["level_1"], 39.0, 5, False, 0.5)
        finally:
            for k, v in old.items():
                if v is None: os.environ.pop(k, None)
                else: os.environ[k] = v
        return out, xs
    ref, xs_ref = run({"CLV_STAGING_MIN_BYTES": str(1 << 40)})           # never staged
    assert ref["level_1"].shape == (2, 11, n, 4)
    per_draw = 2 * n * 4 * 8
    for env in ({"CLV_STAGING_MIN_BYTES": "0"},                                                   # staged, one piece per chain and flush
                {"CLV_STAGING_MIN_BYTES": "0", "CLV_FLUSH_BYTES": str(3 * per_draw)},            # 3 draws per flush
                {"CLV_STAGING_MIN_BYTES": "0", "CLV_FLUSH_BYTES": str(2 * per_draw),
                 "CLV_DRAW_BUFFER_BYTES": str(2 * 5 * per_draw)}):                                # two chunk buffers of 5 draws
        out, xs = run(env)
        for k in ("level_1", "level_2", "loglik_sum"):
            np.testing.assert_array_equal(out[k], ref[k], err_msg=f"{env} {k}")
        np.testing.assert_array_equal(xs, xs_ref, err_msg=str(env))


def test_resident_forecast_equals_host_path(cdnow_abe):
    """Forecast straight from the draws left in HBM == forecast of the same draws through the host API
    (same Philox counters: global customer id, chain-major draw index), and its fused reductions are exact."""
    from mcmc_clv_model_b200.api import _forecast
    d = cdnow_abe
    n = 1500
    with Sampler(d["x"][:n], d["t_x"][:n], d["T_cal"][:n], np.ones((n, 1)), chains=2, seed=3) as s:
        out = s.run(20, 6, 2)
        fr = s.forecast_resident(T_star=39.0, seed=11, want_x_star=True)
        s.run_resident(0, 6, 2)
        fr2 = s.forecast_resident(T_star=39.0, seed=11, want_x_star=True)
    x, _ = _forecast(d["T_cal"][:n], list(out["level_1"]), 39.0, 11, False, 0.5)
    np.testing.assert_array_equal(fr["x_star"], x)
    np.testing.assert_allclose(fr["mean_x_star"], x.mean(axis=0), rtol=1e-13)
    np.testing.assert_allclose(fr["p_alive"], np.concatenate(list(out["level_1"]))[:, :, 3].mean(axis=0), rtol=1e-13)
    assert fr2["x_star"].shape == x.shape and fr2["n_draws_total"] == 6


def test_generator_law_matches_numpy_restatement():
    """Device generator (K5) vs the same law drawn with NumPy: moments of x, t_x, x_star, alive."""
    from mcmc_clv_model_b200.synthetic import C4_BETA, C4_GAMMA, generate_cbs_arrays
    n = 400_000
    g = generate_cbs_arrays(n, C4_BETA, C4_GAMMA, T_cal=(27.0, 38.857), T_star=39.0, seed=123)
    r = np.random.default_rng(1)
    X = np.column_stack([np.ones(n), r.uniform(-1, 1, (n, 4))])
    th = np.exp(X @ C4_BETA + r.multivariate_normal(np.zeros(2), C4_GAMMA, n))
    tau = r.exponential(1 / th[:, 1])
    T = r.uniform(27.0, 38.857, n)
    Te = np.minimum(tau, T)
    x = r.poisson(th[:, 0] * Te)
    tx = np.where(x > 0, Te * r.random(n) ** (1 / np.maximum(x, 1)), 0.0)
    xs = r.poisson(th[:, 0] * np.maximum(0, np.minimum(tau, T + 39.0) - T))
    assert np.all(g["X"][:, 0] == 1) and np.all(np.abs(g["X"][:, 1:]) <= 1)
    assert np.all((g["t_x"] >= 0) & (g["t_x"] <= g["T_cal"])) and np.all((g["x"] == 0) == (g["t_x"] == 0))
    for ours, ref, tol in [(g["x"].mean(), x.mean(), 0.03), ((g["x"] == 0).mean(), (x == 0).mean(), 0.005),
                           (g["t_x"].mean(), tx.mean(), 0.08), (g["x_star"].mean(), xs.mean(), 0.05),
                           ((g["tau_true"] > g["T_cal"]).mean(), (tau > T).mean(), 0.005),
                           (np.log(g["lambda_true"]).mean(), np.log(th[:, 0]).mean(), 0.01),
                           (np.log(g["mu_true"]).std(), np.log(th[:, 1]).std(), 0.01)]:
        assert abs(ours - ref) < tol, (ours, ref)
    # the x | lambda, tau, T law exactly: E[x] = lambda * min(tau, T)
    m = g["lambda_true"] * np.minimum(g["tau_true"], g["T_cal"])
    sel = m < 50
    assert abs((g["x"][sel] - m[sel]).mean()) < 4 * np.sqrt(m[sel].mean() / sel.sum())


def test_api_layout_and_shims(cdnow_abe):
    """The drop-in modules return the reference's dict layout (bi:503-504, tri:653-657) and forecast shapes."""
    import pandas as pd
    import pickle
    from src.models.bivariate.mcmc import draw_future_transactions, mcmc_draw_parameters
    from src.models.trivariate.mcmc import draw_future_transactions as dft3, mcmc_draw_parameters_rfm_m
    d = cdnow_abe
    n = 300
    cbs = pd.DataFrame({k: d[k][:n] for k in ("x", "t_x", "T_cal", "first_sales_scaled", "log_s", "gender_F", "age_scaled")})
    before = cbs.copy()
    draws = mcmc_draw_parameters(cbs, covariates=["first_sales_scaled"], mcmc=21, burnin=10, thin=5, chains=3, seed=1, trace=0)
    pd.testing.assert_frame_equal(cbs, before)                       # caller's frame is never mutated (bi:467)
    assert set(draws) == {"level_1", "level_2", "log_likelihood"} and len(draws["level_1"]) == 3
    assert draws["level_1"][0].shape == (5, n, 4) and draws["level_2"][0].shape == (5, 2 * 2 + 3)
    assert draws["level_1"][0].flags.c_contiguous and draws["level_1"][0].dtype == np.float64
    assert isinstance(draws["log_likelihood"], np.floating) and np.array(draws["level_2"]).shape == (3, 5, 7)
    assert set(np.unique(draws["level_1"][1][:, :, 3])) <= {0.0, 1.0}
    draws2 = pickle.loads(pickle.dumps(draws))
    xs = draw_future_transactions(cbs, draws2, T_star=39.0, seed=4)
    assert xs.shape == (15, n) and xs.dtype == np.int64 and xs.min() >= 0
    t3 = mcmc_draw_parameters_rfm_m(cbs, covariates=["gender_F", "age_scaled"], mcmc=4, burnin=6, thin=1, chains=2, seed=2, trace=0)
    assert t3["level_1"][0].shape == (4, n, 5) and t3["level_2"][0].shape == (4, 3 * 3 + 6) and isinstance(t3["log_likelihood"], float)
    t3["level_1"] = [np.concatenate([a[..., :4], np.log(a[..., 4:])], axis=-1) for a in t3["level_1"]]   # keep exp(eta) finite
    x3, sp = dft3(cbs, t3, T_star=39.0, simulate_spend=True, sigma_s=0.5, seed=9)
    assert x3.shape == sp.shape == (8, n) and np.all((sp > 0) == (x3 > 0))
    assert dft3(cbs, t3, simulate_spend=False, seed=9).shape == (8, n)


def test_screened_poisson_inversion_is_exact_on_adversarial_uniforms():
    """The fp32-screened inversion must return exactly the fp64 CDF-inversion result, also when u sits next to a CDF
    boundary (where the screen has to hand over to the fp64 loop) and for means beyond the fp32 range.  Uniforms are kept
    >= 1e-9 (relative) away from the boundaries: much closer than that the answer depends on the last ulp of exp(-m),
    which CUDA's and NumPy's libm do not share."""
    import ctypes as C
    from mcmc_clv_model_b200 import _lib as L
    rng = np.random.default_rng(11)
    means = np.concatenate([[0.0, 1e-12, 1e-3, 0.5, 1.0, 9.99, 10.0, 59.999, 60.0, 61.0, 150.0, 700.0, 900.0],
                            rng.uniform(0, 70, 400), rng.lognormal(0.0, 1.5, 400)])
    us, ms = [], []
    for m in means:
        k = np.arange(0, int(m + 12 * np.sqrt(m + 1) + 12))
        from scipy.stats import poisson
        cdf = poisson.cdf(k, m)
        picks = cdf[rng.integers(0, len(cdf), 24)]
        cand = np.concatenate([picks * (1 - 1e-9), picks * (1 + 1e-9), picks * (1 - 1e-7), picks * (1 + 1e-7),
                               picks - 3e-6, picks + 3e-6, picks - 1e-4, picks + 1e-4, rng.random(24)])
        cand = np.clip(cand, 1e-300, 1 - 1e-9)     # u within 1e-16 of 1 depends on whether the summed pmf rounds to 1.0
        us.append(cand)
        ms.append(np.full(cand.size, m))
    u = np.concatenate(us)
    m = np.concatenate(ms)
    N = u.size
    l1 = np.zeros((1, N, 4))
    l1[0, :, 0] = m / 39.0
    l1[0, :, 3] = 1.0                                   # alive: horizon = T_star = 39
    xs = np.empty((1, N), dtype=np.int64)
    cfg = L.ForecastConfig(device=0, ncol=4, n_draws_total=1, n_customers=N, T_star=39.0, seed=1)
    L.check(L.load().clv_forecast_injected(C.byref(cfg), L.dptr(l1), L.dptr(np.full(N, 30.0)), L.dptr(np.ascontiguousarray(u[None])),
                                           None, 0, None, xs.ctypes.data_as(L.c_int64_p), None))
    np.testing.assert_array_equal(xs[0], ao.poisson_inversion(l1[0, :, 0] * 39.0, u))


def test_posterior_summary_and_weekly_tracking_on_resident_draws(cdnow_abe):
    """the on-device reductions equal NumPy on the same draws (the reference's
    utils/analysis_bi_helpers.py arithmetic) and the weekly simulation equals its restatement on the same Philox uniforms."""
    from oracle import philox_np as px
    d = cdnow_abe
    n = 400
    with Sampler(d["x"][:n], d["t_x"][:n], d["T_cal"][:n], np.ones((n, 1)), chains=2, seed=3) as s:
        out = s.run(30, 37, 1)
        summ = s.posterior_summary(mu_cap=0.05)
        rng = np.random.default_rng(0)
        birth = rng.uniform(0, 12, n)
        times = np.arange(1.0, 41.0)
        inc = s.weekly_tracking(birth, times, seed=99)
    l1 = np.concatenate(list(out["level_1"]), axis=0)                # (74, n, 4), chain-major like np.concatenate(draws["level_1"])
    np.testing.assert_allclose(summ["mean_lambda"], l1[:, :, 0].mean(axis=0), rtol=1e-12)
    np.testing.assert_allclose(summ["mean_mu_capped"], np.clip(l1[:, :, 1], None, 0.05).mean(axis=0), rtol=1e-12)
    np.testing.assert_allclose(summ["mean_mu"], l1[:, :, 1].mean(axis=0), rtol=1e-12)
    np.testing.assert_allclose(summ["p_alive"], l1[:, :, 3].mean(axis=0), rtol=1e-12)
    np.testing.assert_allclose(summ["mean_tau"], l1[:, :, 2].mean(axis=0), rtol=1e-12)
    for col, key in ((0, "lambda"), (1, "mu")):
        np.testing.assert_allclose(summ[f"{key}_2.5"], np.percentile(l1[:, :, col], 2.5, axis=0), rtol=1e-12)
        np.testing.assert_allclose(summ[f"{key}_97.5"], np.percentile(l1[:, :, col], 97.5, axis=0), rtol=1e-12)
    # weekly tracking: exact restatement with the same counter-based uniforms
    tot = np.zeros(times.size)
    gids = np.arange(n)
    for dd in range(l1.shape[0]):
        lam, tau = l1[dd, :, 0], l1[dd, :, 2]
        for w, t in enumerate(times):
            active = (t > birth) & (t <= birth + tau)                 # analysis_abe.py:456
            u = px.weekly_uniform(99, gids, dd, w)
            tot[w] += ao.poisson_inversion(lam, u)[active].sum()
    np.testing.assert_array_equal(inc, tot / l1.shape[0])
    # and the expectation: sum_i lambda_i * active
    expect = np.array([(l1[:, :, 0] * ((t > birth) & (t <= birth + l1[:, :, 2]))).sum(axis=1).mean() for t in times])
    assert np.all(np.abs(inc - expect) < 6 * np.sqrt(expect / l1.shape[0]) + 0.5)


def test_level1_variates_fast_vs_strict_vs_oracle():
    """The sweep kernel's proposal variates: STRICT == the NumPy restatement of the Philox contract; FAST (fp32 SFU
    transforms, algebraically simplified) == STRICT up to fp32 precision, variate by variate; both are Student-t(3)."""
    import ctypes as C
    from scipy import stats
    from mcmc_clv_model_b200 import _lib as L
    from oracle import philox_np as px
    lib = L.load()
    n, seed, sweep, S = 1_000_000, 777, 5, 5
    v = px.sampler_variates(seed, 0, np.arange(n), sweep, S)
    for step in (0, 1, 4):
        out = {}
        for mode in (L.RNG_STRICT, L.RNG_FAST):
            a, b, u = np.empty(n), np.empty(n), np.empty(n)
            L.check(lib.clv_debug_variates(0, seed, sweep, step, mode, n, L.dptr(a), L.dptr(b), L.dptr(u)))
            out[mode] = (a, b, u)
        np.testing.assert_allclose(out[L.RNG_STRICT][0], v["t3_l"][step], rtol=1e-12)
        np.testing.assert_allclose(out[L.RNG_STRICT][1], v["t3_m"][step], rtol=1e-12)
        np.testing.assert_array_equal(out[L.RNG_STRICT][2], v["u_acc"][step])
        np.testing.assert_array_equal(out[L.RNG_FAST][2], v["u_acc"][step])
        for j in (0, 1):
            s, f = out[L.RNG_STRICT][j], out[L.RNG_FAST][j]
            err = np.abs(f - s) / (1e-3 + np.abs(s))
            assert np.median(err) < 1e-5 and np.quantile(err, 0.999) < 1e-2, (np.median(err), np.quantile(err, 0.999))
            assert stats.kstest(f, stats.t(3).cdf).pvalue > 1e-3
            assert stats.kstest(s, stats.t(3).cdf).pvalue > 1e-3
        # the accept uniform shares its words with the proposals (low bytes vs top 24 bits): no visible dependence
        assert stats.kstest(out[L.RNG_FAST][2], "uniform").pvalue > 1e-3
        for j in (0, 1):
            assert abs(np.corrcoef(out[L.RNG_FAST][2], np.abs(out[L.RNG_FAST][j]) < 1.0)[0, 1]) < 0.01
    assert abs(np.corrcoef(out[L.RNG_FAST][0], out[L.RNG_FAST][1])[0, 1]) < 0.01


@pytest.mark.parametrize("name", ["abe", "full"])
def test_elog2cbs_vs_reference(name):
    """device CBS builder == the reference's pandas `elog2cbs` on the CDNOW event logs (counts and
    day arithmetic exact, sums to 1e-12), events fed in shuffled order."""
    import pandas as pd
    from mcmc_clv_model_b200.cbs import elog2cbs
    g = load_golden(f"elog_{name}.npz")
    rng = np.random.default_rng(0)
    perm = rng.permutation(g["cust"].size)
    elog = pd.DataFrame({"cust": g["cust"][perm], "date": pd.Timestamp("1970-01-01") + pd.to_timedelta(g["day"][perm].astype(np.int64), unit="D"),
                         "sales": g["sales"][perm]})
    cbs = elog2cbs(elog, units="W", T_cal="1997-09-30", T_tot="1998-06-30")
    assert list(cbs.columns) == ["cust", "x", "t_x", "litt", "sales", "sales_x", "first", "T_cal", "T_star", "x_star", "sales_star"]
    np.testing.assert_array_equal(cbs["cust"].to_numpy(), g["cbs_cust"])
    np.testing.assert_array_equal(cbs["x"].to_numpy(), g["cbs_x"])
    np.testing.assert_array_equal(cbs["x_star"].to_numpy(), g["cbs_x_star"])
    np.testing.assert_array_equal(((cbs["first"] - pd.Timestamp("1970-01-01")) // pd.Timedelta(days=1)).to_numpy(), g["cbs_first"])
    for c in ("t_x", "T_cal", "T_star"):
        np.testing.assert_allclose(cbs[c].to_numpy(), g[f"cbs_{c}"], rtol=1e-14, atol=1e-14)
    for c in ("litt", "sales", "sales_x", "sales_star"):
        np.testing.assert_allclose(cbs[c].to_numpy(), g[f"cbs_{c}"], rtol=1e-12, atol=1e-10)
    # no hold-out, no sales column
    c2 = elog2cbs(elog[["cust", "date"]], units="D")
    assert list(c2.columns) == ["cust", "x", "t_x", "litt", "sales", "sales_x", "first", "T_cal"] and (c2["sales"] >= c2["x"] + 1).all()   # same-day events merge: sales counts events


def test_analysis_helpers_on_a_draws_dict(cdnow_abe):
    """The analysis reductions on an (unpickled-style) draws dict: upload once, reduce on the device; Table-4 columns
    against the reference's formulas in NumPy (utils/analysis_bi_helpers.py:75-110)."""
    import pandas as pd
    from mcmc_clv_model_b200 import mcmc_draw_parameters
    from mcmc_clv_model_b200.analysis import posterior_summary, table4_inputs, weekly_tracking
    d = cdnow_abe
    n = 250
    cbs = pd.DataFrame({k: d[k][:n] for k in ("x", "t_x", "T_cal")})
    draws = mcmc_draw_parameters(cbs, mcmc=40, burnin=30, thin=2, chains=3, seed=8, trace=0)
    allv = np.concatenate(draws["level_1"], axis=0)
    s = posterior_summary(draws)
    np.testing.assert_allclose(s["mean_lambda"], allv[:, :, 0].mean(axis=0), rtol=1e-12)
    np.testing.assert_allclose(s["mu_97.5"], np.percentile(allv[:, :, 1], 97.5, axis=0), rtol=1e-12)
    t4 = table4_inputs(draws)
    mean_mu = np.clip(allv[:, :, 1], None, 0.05).mean(axis=0)
    mean_z = allv[:, :, 3].mean(axis=0)
    np.testing.assert_allclose(t4["Exp # of trans in val period"],
                               mean_z * (allv[:, :, 0].mean(axis=0) / mean_mu) * (1 - np.exp(-mean_mu * 39)), rtol=1e-10)
    inc = weekly_tracking(draws, np.zeros(n), np.arange(1.0, 30.0), seed=1)
    expect = np.array([(allv[:, :, 0] * (t <= allv[:, :, 2])).sum(axis=1).mean() for t in np.arange(1.0, 30.0)])
    assert np.all(np.abs(inc - expect) < 6 * np.sqrt(expect / allv.shape[0]) + 0.5)


def test_exported_draw_z_and_draw_tau_blocks(cdnow_abe):
    """`draw_z` / `draw_tau` of the reference's `__all__` (bi:193-227): same signature, same consumption of the NumPy
    generator, arithmetic on the device; z bit-exact and tau to 1e-12 against the oracle driven by the same generator."""
    import pandas as pd
    from src.models.bivariate.mcmc import draw_tau, draw_z
    d = cdnow_abe
    n = 600
    cbs = pd.DataFrame({k: d[k][:n] for k in ("x", "t_x", "T_cal")})
    g = np.random.default_rng(3)
    lam, mu = g.lognormal(-3.3, 1.0, n), g.lognormal(-3.5, 1.2, n)
    z = draw_z(cbs, lam, mu, np.random.default_rng(10))
    z_ref = ao.draw_z(d["t_x"][:n], d["T_cal"][:n], lam, mu, np.random.default_rng(10).random(n))
    np.testing.assert_array_equal(z, z_ref)
    tau = draw_tau(cbs, lam, mu, z, np.random.default_rng(11))
    r = np.random.default_rng(11)
    e, u = np.zeros(n), np.zeros(n)
    e[z] = r.standard_exponential(int(z.sum()))
    u[~z] = r.random(int((~z).sum()))
    np.testing.assert_allclose(tau, ao.draw_tau(d["t_x"][:n], d["T_cal"][:n], lam, mu, z, e, u), rtol=1e-12)
    assert np.all(tau[z] > d["T_cal"][:n][z])


@pytest.mark.parametrize("D", [2, 3])
def test_checkpoint_resume_is_bit_identical(cdnow_abe, D):
    """The final state + sweep counter continue a chain in a fresh handle exactly (counter-based RNG: nothing else to save)."""
    d = cdnow_abe
    n = 900
    X = np.column_stack([np.ones(n), d["age_scaled"][:n]])
    args = (d["x"][:n], d["t_x"][:n], d["T_cal"][:n], X, d["log_s"][:n] if D == 3 else None)
    kw = dict(model_dim=D, chains=2, seed=31)
    with Sampler(*args, **kw) as s:
        s.run(0, 9, 3)
        ck = s.checkpoint()
        ref = s.run(2, 6, 2)
    with Sampler(*args, **kw) as s2:
        s2.restore(ck)
        out = s2.run(2, 6, 2)
    for k in ("level_1", "level_2", "loglik_sum"):
        np.testing.assert_array_equal(out[k], ref[k], err_msg=k)


@pytest.mark.parametrize("D,K", [(2, 9), (3, 16)])
def test_many_covariates_vs_oracle(D, K):
    """The upper end of the design-matrix width (CLV_MAX_K = 16): level-2 lane loops beyond one warp's width, the
    statistic columns beyond 48 KB of shared memory, stream and persistent paths -- against the oracle (strict Philox)."""
    rng = np.random.default_rng(K)
    n = 517
    x = rng.poisson(1.3, n)
    T = rng.uniform(27, 39, n)
    t_x = np.where(x > 0, T * rng.random(n), 0.0)
    X = np.column_stack([np.ones(n), rng.normal(size=(n, K - 1)) * 0.7])
    log_s = rng.normal(3.0, 0.5, n) if D == 3 else None
    cbs = ao.Cbs(x=x.astype(np.int64), t_x=t_x, T_cal=T, X=X, log_s=log_s)
    ora = ao.run_chain(cbs, ao.default_hyper(K, D), PhiloxStreams(5, 0, np.arange(n), 6, D, K), mcmc=3, burnin=1, thin=1, D=D,
                       n_mh_steps=6)
    for mode in ("stream", "persistent"):
        with Sampler(x, t_x, T, X, log_s, model_dim=D, chains=1, n_mh_steps=6, seed=5, rng="strict", sweep_mode=mode) as s:
            out = s.run(1, 3, 1)
        np.testing.assert_array_equal(out["level_1"][0][:, :, 3], ora["level_1"][:, :, 3])
        np.testing.assert_allclose(out["level_1"][0], ora["level_1"], rtol=RTOL)
        np.testing.assert_allclose(out["level_2"][0], ora["level_2"], rtol=RTOL, atol=1e-9)


def test_compat_paper_beta_draw_vs_oracle():
    """compat="paper": beta | Sigma ~ matrix-normal(B_hat, V, Sigma) as Abe (2009) writes it (B_hat + chol(V) Z chol(Sigma)'),
    instead of the reference's kron-ordered noise (bi:261); same injected streams through the oracle."""
    g = load_golden("inj_bi_k4.npz")
    D, S = int(g["D"]), int(g["S"])
    cbs = ao.Cbs(x=g["x"], t_x=g["t_x"], T_cal=g["T_cal"], X=g["X"])
    T = g["u_z"].shape[0]
    ora = ao.run_chain(cbs, ao.default_hyper(cbs.K, D), ReplayStreams(g), mcmc=T, burnin=0, thin=1, D=D, n_mh_steps=S, compat="paper")
    assert np.abs(ora["level_2"] - g["level_2"]).max() > 1e-3            # the two conventions do differ for K > 1
    with Sampler(g["x"], g["t_x"], g["T_cal"], g["X"], model_dim=D, chains=1, n_mh_steps=S, rng="injected", compat="paper") as s:
        for t in range(T):
            out = s.sweep_injected(_sweep_arrays(g, t, D))
            np.testing.assert_allclose(out["level_2"][0], ora["level_2"][t], rtol=RTOL, atol=1e-9)
            np.testing.assert_array_equal(out["level_1"][0][:, 3], ora["level_1"][t][:, 3])


def _device_fast_variates(seed, n, S):
    """callable(sweep) -> the FAST-mode Metropolis variates of chain 0 exactly as k_sweep<.,FAST> consumes them
    (clv_debug_variates runs the kernel's own t3_fast / low_bytes code)."""
    from mcmc_clv_model_b200 import _lib as L
    lib = L.load()

    def variates(sweep):
        out = {k: np.empty((S, n)) for k in ("t3_l", "t3_m", "u_acc")}
        for step in range(S):
            L.check(lib.clv_debug_variates(0, seed, sweep, step, L.RNG_FAST, n, L.dptr(out["t3_l"][step]),
                                           L.dptr(out["t3_m"][step]), L.dptr(out["u_acc"][step])))
        return out
    return variates


@pytest.mark.parametrize("mode", ["stream", "stream2", "persistent"])
@pytest.mark.parametrize("D,cov", [(2, ["first_sales_scaled"]), (3, ["gender_F", "age_scaled"])])
def test_fast_kernel_trajectory_vs_oracle_at_full_cdnow_size(cdnow_full, D, cov, mode, monkeypatch):
    """The TIMED kernel (rng="fast": fp32 SFU proposal variates, fp32-screened accept) at the size of BASELINE.json
    configs[1] / configs[2] (23 570 customers = 185 tiles; C2: K=2 bivariate, C3: K=3 trivariate): its FAST variates
    are pulled from the device and replayed through the oracle, and the whole production trajectory (z, tau, 20
    Metropolis steps, eta, level-2, burn-in / thinning) must match: z bit-exact, continuous 1e-6 (north_star)."""
    d = cdnow_full
    n = d["x"].size
    X = np.column_stack([np.ones(n)] + [d[c].astype(float) for c in cov])
    cbs = ao.Cbs(x=d["x"].astype(np.int64), t_x=d["t_x"], T_cal=d["T_cal"], X=X, log_s=d["log_s"] if D == 3 else None)
    seed, burnin, mcmc, thin, S = 2025, 1, 3, 1, 20
    ora = ao.run_chain(cbs, ao.default_hyper(cbs.K, D),
                       PhiloxStreams(seed, 0, np.arange(n), S, D, cbs.K, level1_variates=_device_fast_variates(seed, n, S)),
                       mcmc=mcmc, burnin=burnin, thin=thin, D=D, n_mh_steps=S)
    # "stream2": the sweep kernel with two customers per thread (k_sweep2; the default once customers x chains >= 100 000,
    # i.e. the kernel bench.py times), forced here; "stream": one customer per thread (the default at this size)
    monkeypatch.setenv("CLV_SWEEP_CPT", "2" if mode == "stream2" else "1")
    with Sampler(cbs.x, cbs.t_x, cbs.T_cal, X, cbs.log_s, model_dim=D, chains=1, n_mh_steps=S, seed=seed, rng="fast",
                 sweep_mode="stream" if mode == "stream2" else mode) as s:
        out = s.run(burnin, mcmc, thin)
    np.testing.assert_array_equal(out["level_1"][0][:, :, 3], ora["level_1"][:, :, 3])
    np.testing.assert_allclose(out["level_1"][0], ora["level_1"], rtol=RTOL)
    np.testing.assert_allclose(out["level_2"][0], ora["level_2"], rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(out["loglik_sum"][0] / n, ora["log_likelihood"], rtol=RTOL)


def test_one_and_two_customers_per_thread_give_the_same_chain(cdnow_full, monkeypatch):
    """k_sweep (one customer per thread) and k_sweep2 (two: the variant that runs when the problem fills the GPU) do the
    same arithmetic per customer: 2 chains x 23 570 customers, strict and fast, trivariate too -- bit-identical draws."""
    d = cdnow_full
    n = d["x"].size
    X = np.column_stack([np.ones(n), d["first_sales_scaled"], d["age_scaled"]])
    for D, rng in ((2, "fast"), (2, "strict"), (3, "fast")):
        res = []
        for cpt in ("1", "2"):
            monkeypatch.setenv("CLV_SWEEP_CPT", cpt)
            with Sampler(d["x"], d["t_x"], d["T_cal"], X, d["log_s"] if D == 3 else None, model_dim=D, chains=2, seed=5, rng=rng,
                         sweep_mode="stream") as s:
                res.append(s.run(2, 4, 2))
        for k in ("level_1", "level_2", "loglik_sum"):
            np.testing.assert_array_equal(res[0][k], res[1][k], err_msg=f"{k} D={D} rng={rng}")


def test_strict_trajectory_spanning_many_tiles_and_blocks(cdnow_full):
    """STRICT Philox at 23 570 customers with 2 chains in one handle (the chain index travels in the counter; 185 tiles
    per chain, i.e. far more than the single tile of the injected goldens)."""
    d = cdnow_full
    n = d["x"].size
    X = np.column_stack([np.ones(n), d["first_sales_scaled"]])
    cbs = ao.Cbs(x=d["x"].astype(np.int64), t_x=d["t_x"], T_cal=d["T_cal"], X=X)
    seed, S = 99, 20
    with Sampler(cbs.x, cbs.t_x, cbs.T_cal, X, model_dim=2, chains=2, n_mh_steps=S, seed=seed, rng="strict", sweep_mode="stream") as s:
        out = s.run(1, 2, 1)
    for chain in (0, 1):
        ora = ao.run_chain(cbs, ao.default_hyper(2, 2), PhiloxStreams(seed, chain, np.arange(n), S, 2, 2), mcmc=2, burnin=1,
                           thin=1, D=2, n_mh_steps=S)
        np.testing.assert_array_equal(out["level_1"][chain][:, :, 3], ora["level_1"][:, :, 3])
        np.testing.assert_allclose(out["level_1"][chain], ora["level_1"], rtol=RTOL)
        np.testing.assert_allclose(out["level_2"][chain], ora["level_2"], rtol=RTOL, atol=1e-9)


def test_generator_matches_the_reference_generator_moments():
    """`generate_pareto_abe` through the drop-in module vs golden moments of the reference's OWN generator
    (tests/golden/make_generator_golden.py ran bi:95-187 unmodified: n = 20 000, scalar and vector T_cal).  Every
    statistic within 4 combined standard errors; and the reference's CBS convention for cohorts (bi:158,165): t_x on
    the shifted clock, one scalar T_cal."""
    from mcmc_clv_model_b200.synthetic import generate_pareto_abe
    g = load_golden("gen_ref.npz")
    names = [str(s) for s in g["names"]]
    n, T_star = int(g["n"]), float(g["T_star"])

    def moments(cbs, T_rel, t_rel):
        x, xs = cbs["x"].to_numpy(float), cbs["x_star"].to_numpy(float)
        alive = cbs["alive_true"].to_numpy(float)
        st = {"p_x0": x == 0, "p_x_ge1": x >= 1, "p_x_ge3": x >= 3, "p_x_ge10": x >= 10, "mean_sqrt_x": np.sqrt(x),
              "mean_log1p_x": np.log1p(x), "p_xs0": xs == 0, "p_xs_ge3": xs >= 3, "mean_log1p_xs": np.log1p(xs), "alive": alive,
              "mean_tx_over_T": t_rel / T_rel, "p_tx_late": (t_rel / T_rel) > 0.75,
              "mean_log_lambda": np.log(cbs["lambda_true"].to_numpy()), "mean_log_mu": np.log(cbs["mu_true"].to_numpy()),
              "corr_proxy": np.log1p(x) * alive}
        val = np.array([np.mean(st[k].astype(float)) for k in names])
        se = np.array([np.std(st[k].astype(float), ddof=1) / np.sqrt(len(x)) for k in names])
        return val, se

    # scalar T_cal: ten times the reference's sample, so the error is the golden's
    big = 10 * n
    cbs, elog = generate_pareto_abe(big, 32.0, T_star, g["beta"], g["gamma"], seed=42)
    assert list(cbs.columns[:4]) == ["cust", "x", "t_x", "T_cal"] and np.all(cbs["T_cal"] == 32.0)
    val, se = moments(cbs, np.full(big, 32.0), cbs["t_x"].to_numpy())
    z = np.abs(val - g["scalar_val"]) / np.hypot(se, g["scalar_se"])
    assert z.max() < 4.0, dict(zip(names, np.round(z, 2)))
    # vector T_cal with the golden's covariate and T_cal inputs: law and CBS convention
    T_vec, T_fix = g["vector_T_cal_in"], float(g["vector_T_fix"])
    cbs2, elog2 = generate_pareto_abe(n, T_vec, T_star, g["beta"], g["gamma"], covars=g["vector_cov"], seed=43)
    T_zero = T_fix - T_vec
    assert np.all(cbs2["T_cal"].to_numpy() == T_fix)                                   # bi:165: one scalar T_cal
    t_rel = cbs2["t_x"].to_numpy() - T_zero
    assert np.all(t_rel >= 0) and np.all((cbs2["x"].to_numpy() == 0) == (t_rel == 0))   # x = 0 => t_x = T_zero (first purchase)
    np.testing.assert_allclose(elog2.groupby("cust")["t"].min().to_numpy(), T_zero)     # event log on the shifted clock
    val2, se2 = moments(cbs2, T_vec, t_rel)
    z2 = np.abs(val2 - g["vector_val"]) / np.hypot(se2, g["vector_se"])
    assert z2.max() < 4.0, dict(zip(names, np.round(z2, 2)))
    # the customer clock is still available, and equals the reference clock for scalar T_cal
    cbs3, _ = generate_pareto_abe(n, T_vec, T_star, g["beta"], g["gamma"], covars=g["vector_cov"], seed=43, cbs_clock="customer")
    np.testing.assert_allclose(cbs3["t_x"].to_numpy(), t_rel, atol=1e-12)
    np.testing.assert_array_equal(cbs3["T_cal"].to_numpy(), T_vec)


@pytest.mark.parametrize("D,rng,G", [(2, "fast", 3), (2, "strict", 2), (3, "fast", 4)])
def test_customer_shards_in_lockstep_equal_the_unsharded_chain(D, rng, G):
    """The customer-sharded path on ONE GPU (the driver's test box has one): G shards of one problem -- contiguous
    1024-aligned customer ranges, Philox counters on global ids, int64 fixed-point level-2 statistics -- advance in
    lockstep, the per-sweep all-reduce running the production peer-mailbox (LL) protocol with the ranks emulated as the
    blocks of one cooperative launch.  Every shard must carry exactly the unsharded chain: the property behind
    "N-GPU == 1-GPU" (tests/test_gpu_multi.py checks the real transports on >= 2 GPUs)."""
    from mcmc_clv_model_b200.distributed import shard_bounds
    from mcmc_clv_model_b200.synthetic import C4_BETA, C4_GAMMA, C4_SEED, C4_T_CAL, generate_cbs_arrays
    N, chains, sweeps = 40_003, 2, 4
    c = generate_cbs_arrays(N, C4_BETA, C4_GAMMA, T_cal=C4_T_CAL, seed=C4_SEED, with_truth=False)
    log_s = (0.5 * c["X"][:, 1] + 3.0 + 0.1 * np.cos(np.arange(N))) if D == 3 else None
    with Sampler(c["x"], c["t_x"], c["T_cal"], c["X"], log_s, model_dim=D, chains=chains, seed=9, rng=rng, sweep_mode="stream") as full:
        stats = dict(full.init_stats)
        full.run(sweeps - 1, 1, 1, store_level1=False)          # (clv_run keeps z / tau of its last sweep; clv_advance does not)
        ref = [full.get_state(ch) for ch in range(chains)]
    shards = []
    try:
        for lo, hi in shard_bounds(N, G):
            shards.append(Sampler(c["x"][lo:hi], c["t_x"][lo:hi], c["T_cal"][lo:hi], c["X"][lo:hi], None if log_s is None else log_s[lo:hi],
                                  model_dim=D, chains=chains, seed=9, rng=rng, sweep_mode="stream", n_global=N, gid_offset=lo,
                                  init_stats=stats))
        Sampler.lockstep_advance(shards, sweeps)
        for (lo, hi), s in zip(shard_bounds(N, G), shards):
            assert s.sweeps_done == sweeps
            for ch in range(chains):
                st = s.get_state(ch)
                np.testing.assert_array_equal(st["beta"], ref[ch]["beta"])
                np.testing.assert_array_equal(st["Sigma"], ref[ch]["Sigma"])
                for k in ("log_lambda", "log_mu", "z", "tau") + (("log_eta",) if D == 3 else ()):
                    np.testing.assert_array_equal(st[k], ref[ch][k][lo:hi], err_msg=f"{k} shard [{lo},{hi}) chain {ch}")
    finally:
        for s in shards:
            s.close()


def test_covariate_standardisation_vs_the_committed_full_cbs():
    """SURVEY 8f row f-3, second half: first_sales_scaled / age_scaled / gender_binary of
    src/data_processing/2B_cdnow_elog2cbs_full.py:62-105 computed on the device from the raw event log and customer table,
    against the columns committed in data/processed/cdnow_fullCBS.csv (tests/golden/covar_full.npz)."""
    import pandas as pd
    from mcmc_clv_model_b200.cbs import add_covariates, elog2cbs, standardize
    g, c = load_golden("elog_full.npz"), load_golden("covar_full.npz")
    rng = np.random.default_rng(1)
    elog = pd.DataFrame({"cust": g["cust"], "date": pd.Timestamp("1970-01-01") + pd.to_timedelta(g["day"].astype(np.int64), unit="D"),
                         "sales": g["sales"]})
    cbs = elog2cbs(elog, units="W", T_cal="1997-09-30", T_tot="1998-06-30", with_first_sales=True)
    # groupby.first semantics: the first row of each customer in INPUT order
    np.testing.assert_array_equal(cbs["first_sales"].to_numpy(), elog.groupby("cust")["sales"].first().to_numpy())
    customers = pd.DataFrame({"cust": c["cust"], "age": c["age"], "gender": np.where(c["gender_is_M"] == 1, "M", "F"),
                              "zone": "z", "state": "s", "age_category": "a"}).sample(frac=1.0, random_state=3)      # any row order
    full = add_covariates(cbs, elog, customers)
    assert not {"gender", "zone", "state", "age_category", "first_sales"} & set(full.columns)
    np.testing.assert_array_equal(full["cust"].to_numpy(), c["cust"])
    np.testing.assert_allclose(full["first_sales_scaled"].to_numpy(), c["first_sales_scaled"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(full["age_scaled"].to_numpy(), c["age_scaled"], rtol=1e-12, atol=1e-13)
    np.testing.assert_array_equal(full["gender_binary"].to_numpy(), c["gender_binary"])
    # deterministic, and equal to pandas on a hostile column (large offset, shuffled)
    v = rng.normal(1e6, 3.0, 200_001)
    z1, m1, s1 = standardize(v)
    z2, m2, s2 = standardize(v)
    assert np.array_equal(z1, z2) and (m1, s1) == (m2, s2)
    ser = pd.Series(v)
    assert abs(m1 - ser.mean()) < 1e-9 and abs(s1 / ser.std() - 1) < 1e-12
    np.testing.assert_allclose(z1, ((ser - ser.mean()) / ser.std()).to_numpy(), atol=1e-9)
    # first row in input order is not the earliest date once the log is shuffled
    perm = rng.permutation(len(elog))
    sh = elog.iloc[perm].reset_index(drop=True)
    cb2 = elog2cbs(sh, units="W", T_cal="1997-09-30", T_tot="1998-06-30", with_first_sales=True)
    np.testing.assert_array_equal(cb2["first_sales"].to_numpy(), sh.groupby("cust")["sales"].first().to_numpy())


def test_forecast_spend_philox_vs_restated_contract():
    """Production trivariate forecast with spend (tri:722-741, Q7): per-transaction normals from the forecast domain's
    spend slots (disjoint from the Poisson / PTRS / weekly-tracking slots) against the NumPy restatement, and the bounded
    branch for huge counts (normal limit of the sum of log-normals) against its law."""
    from oracle import philox_np as px
    from mcmc_clv_model_b200.api import _forecast
    g = load_golden("fc_tri.npz")
    l1 = g["level_1"][:2].copy()
    nd, N, _ = l1.shape
    l1[:, :, 4] = np.clip(l1[:, :, 4], 0.5, 4.0)                 # eta is used as the log-mean (Q7): keep exp() tame
    l1[:, 3, 0], l1[:, 3, 3] = 200.0 / 39.0, 1.0                 # one customer through the PTRS branch, > 126 transactions:
    x, sp = _forecast(g["T_cal"], [l1], 39.0, 31, True, 0.5)     # its spend normals used to share slots with the PTRS attempts
    ref = np.zeros((nd, N))
    for d in range(nd):
        for i in range(N):
            ref[d, i] = sum(np.exp(l1[d, i, 4] + 0.5 * px.forecast_spend_normal(31, i, d, j)) for j in range(int(x[d, i])))
    assert x[:, 3].min() > 126
    np.testing.assert_allclose(sp, ref, rtol=1e-12, atol=1e-12)
    # huge counts: bounded work, right mean and spread
    M = 4000
    big = np.zeros((2, M, 5))
    big[:, :, 0], big[:, :, 3], big[:, :, 4] = 20000.0 / 39.0, 1.0, 1.0
    xb, sb = _forecast(np.full(M, 30.0), [big], 39.0, 7, True, 0.5)
    assert xb.min() > 4096
    m1, v1 = np.exp(1.0 + 0.125), (np.exp(0.25) - 1.0) * np.exp(2.0 + 0.25)
    zs = (sb - xb * m1) / np.sqrt(xb * v1)
    assert abs(zs.mean()) < 5 / np.sqrt(zs.size) and abs(zs.std() - 1) < 0.05


@pytest.mark.parametrize("D", [2, 3])
def test_fused_forecast_equals_forecast_of_the_stored_draws(cdnow_abe, D):
    """x* drawn inside the sweep kernel when a draw is kept (lambda, tau, z in registers: zero re-read) == the forecast
    of the same stored draws with the same seed (clv_forecast_resident, and the host-path clv_forecast): same Philox
    counters (global customer id, draw index chain * n_draws + draw).  Includes a customer in the PTRS regime."""
    from mcmc_clv_model_b200.api import _forecast
    d = cdnow_abe
    n = 1500
    x = d["x"][:n].copy()
    x[7] = 900                                                       # lambda * T_star well beyond the PTRS switch
    t_x = d["t_x"][:n].copy()
    t_x[7] = d["T_cal"][7] - 0.01
    X = np.column_stack([np.ones(n), d["first_sales_scaled"][:n]])
    with Sampler(x, t_x, d["T_cal"][:n], X, d["log_s"][:n] if D == 3 else None, model_dim=D, chains=3, seed=21) as s:
        s.set_fused_forecast(T_star=39.0, seed=77)
        out = s.run(40, 30, 3)                                        # 10 kept draws per chain, draws stored as well
        fused = s.fused_forecast_result()
        res = s.forecast_resident(T_star=39.0, seed=77, want_x_star=True)
        s.set_fused_forecast(T_star=39.0, seed=78)
        s.run(0, 12, 4, store_level1=False)                           # draws not stored at all
        fused2 = s.fused_forecast_result()
        s.set_fused_forecast(enable=False)
    assert fused["n_draws_total"] == 30 and fused2["n_draws_total"] == 9
    np.testing.assert_array_equal(fused["mean_x_star"], res["mean_x_star"])
    np.testing.assert_array_equal(fused["p_alive"], res["p_alive"])
    xs, _ = _forecast(d["T_cal"][:n], list(out["level_1"]), 39.0, 77, False, 0.5)
    np.testing.assert_allclose(fused["mean_x_star"], xs.mean(axis=0), rtol=1e-13)
    assert res["x_star"][:, 7].min() > 60 and np.isfinite(fused2["mean_x_star"]).all()


def test_resident_forecast_list_overflow_is_repeated_with_a_larger_list(cdnow_abe, monkeypatch):
    """The two-pass resident forecast lists the cells that need more than the quick path; when the list is too small the
    pass is repeated with one sized for the count (CLV_FC_LIST_CAP forces that here) -- same result."""
    d = cdnow_abe
    n = 2000
    with Sampler(d["x"][:n], d["t_x"][:n], d["T_cal"][:n], np.ones((n, 1)), chains=2, seed=3) as s:
        s.run_resident(30, 40, 1)
        ref = s.forecast_resident(T_star=39.0, seed=9, want_x_star=True)
        monkeypatch.setenv("CLV_FC_LIST_CAP", "7")
        small = s.forecast_resident(T_star=39.0, seed=9, want_x_star=True)
        monkeypatch.delenv("CLV_FC_LIST_CAP")
        monkeypatch.setenv("CLV_FC_KERNEL", "tma")          # the TMA-fed main pass (cp.async.bulk ring): same x*
        again = s.forecast_resident(T_star=39.0, seed=9, want_x_star=True)
        monkeypatch.setenv("CLV_FC_LIST_CAP", "7")
        again_small = s.forecast_resident(T_star=39.0, seed=9, want_x_star=True)
    for k in ("mean_x_star", "p_alive", "x_star"):
        np.testing.assert_array_equal(small[k], ref[k])
        np.testing.assert_array_equal(again[k], ref[k])
        np.testing.assert_array_equal(again_small[k], ref[k])
    assert ref["x_star"].max() >= 8                          # some cells did take the second pass


@pytest.mark.parametrize("D", [2, 3])
def test_resident_forecast_in_chunks_equals_one_pass(cdnow_abe, monkeypatch, D):
    """The resident forecast can run in chunks of draw pairs, the second pass of a chunk on a side stream beside the main
    pass of the next (CLV_FC_CHUNKS; an A/B knob, measured no faster): same x*, same sums, for an odd number of draws,
    with and without x* written out, and when a chunk's list overflows (repeated in one pass)."""
    d = cdnow_abe
    n = 2357
    log_s = d["log_s"][:n] if D == 3 else None
    with Sampler(d["x"][:n], d["t_x"][:n], d["T_cal"][:n], np.ones((n, 1)), log_s, model_dim=D, chains=3, seed=4) as s:
        s.run_resident(30, 37, 1)                             # 3 x 37 = 111 draws: 56 pairs, the last one half empty
        monkeypatch.setenv("CLV_FC_CHUNKS", "1")
        ref = s.forecast_resident(T_star=39.0, seed=9, want_x_star=True)
        ref_sums = s.forecast_resident(T_star=39.0, seed=9)
        outs = []
        for chunks, side, cap in (("2", "1", None), ("5", "2", None), ("56", "1", None), ("7", "1", "70")):
            monkeypatch.setenv("CLV_FC_CHUNKS", chunks)
            monkeypatch.setenv("CLV_FC_SIDE_BLOCKS", side)
            if cap: monkeypatch.setenv("CLV_FC_LIST_CAP", cap)
            outs.append((s.forecast_resident(T_star=39.0, seed=9, want_x_star=True), s.forecast_resident(T_star=39.0, seed=9)))
    for full, sums in outs:
        for k in ("mean_x_star", "p_alive", "x_star"):
            np.testing.assert_array_equal(full[k], ref[k], err_msg=k)
        for k in ("mean_x_star", "p_alive"):
            np.testing.assert_array_equal(sums[k], ref_sums[k], err_msg=k)
            np.testing.assert_array_equal(sums[k], ref[k], err_msg=k)
    assert ref["x_star"].max() >= 8


@pytest.mark.parametrize("D,cov", [(2, ["first_sales_scaled"]), (3, ["gender_F", "age_scaled"])])
def test_injected_sweeps_at_full_cdnow_size_vs_oracle(cdnow_full, D, cov):
    """Injected streams at the size of BASELINE.json configs[1] / [2] (23 570 customers, 185 tiles -- the reference-made
    injected goldens hold a single tile): the CUDA sweep fed seeded variates against the oracle fed the same arrays
    (the oracle is pinned to the reference's own `_run_chain` by tests/test_oracle_golden.py).  z bit-exact, continuous
    1e-6, three sweeps, the statistics of the second and third ones reduced over all tiles by the kernel."""
    from oracle.streams import random_replay_arrays
    d = cdnow_full
    n = d["x"].size
    X = np.column_stack([np.ones(n)] + [d[c].astype(float) for c in cov])
    cbs = ao.Cbs(x=d["x"].astype(np.int64), t_x=d["t_x"], T_cal=d["T_cal"], X=X, log_s=d["log_s"] if D == 3 else None)
    S, T = 6, 3
    hy = ao.default_hyper(cbs.K, D)
    arrays = random_replay_arrays(np.random.default_rng(11 + D), T, n, S, D, cbs.K, hy["nu_00"] + n)
    ora = ao.run_chain(cbs, hy, ReplayStreams(arrays), mcmc=T, burnin=0, thin=1, D=D, n_mh_steps=S)
    with Sampler(cbs.x, cbs.t_x, cbs.T_cal, X, cbs.log_s, model_dim=D, chains=1, n_mh_steps=S, rng="injected") as s:
        for t in range(T):
            out = s.sweep_injected({k: v[t][None] for k, v in arrays.items()}, keep=True)
            np.testing.assert_array_equal(out["level_1"][0][:, 3], ora["level_1"][t][:, 3], err_msg=f"z differs at sweep {t}")
            np.testing.assert_allclose(out["level_1"][0], ora["level_1"][t], rtol=RTOL, err_msg=f"level_1 sweep {t}")
            np.testing.assert_allclose(out["level_2"][0], ora["level_2"][t], rtol=RTOL, atol=1e-9, err_msg=f"level_2 sweep {t}")
            np.testing.assert_allclose(out["loglik_sum"][0] / n, ora["log_likelihood"][t], rtol=RTOL)


@pytest.mark.parametrize("cpt", ["1", "2"])
@pytest.mark.parametrize("case", ["regular_sigma", "near_singular_sigma"])
def test_fast_kernel_exact_path_and_guards_on_adversarial_rows(cpt, case, monkeypatch):
    """The fp32-screened Metropolis decision (CLV_E32) must ALWAYS equal the fp64 expression of the reference: rows that
    fail the screen's rounding guard (x = 6e7 transactions; a nearly singular Sigma, where every customer fails it),
    proposals beyond the +-70 clip, log mu crossing 5 (target -inf, bi:308-309) and states that start there.  Trivariate
    order (level 1 first, so the Sigma given to set_state is the one the Metropolis steps see); the kernel's own FAST
    variates are replayed through the oracle from the same state: z bit-exact, continuous 1e-6, after each of 3 sweeps."""
    rs = np.random.RandomState(11)
    n, S, D, seed = 700, 20, 3, 77
    T = np.full(n, 39.0)
    x = rs.poisson(3.0, n).astype(np.int64)
    t_x = np.where(x > 0, rs.uniform(1.0, 38.0, n), 0.0)
    x[:10] = 60_000_000                       # 420 x > the guard's budget: always the exact path
    t_x[:10] = 35.0
    X = np.column_stack([np.ones(n), rs.normal(size=n)])
    log_s = rs.normal(3.5, 0.6, n)
    cbs = ao.Cbs(x=x, t_x=t_x, T_cal=T, X=X, log_s=log_s)
    hyper = ao.default_hyper(2, D)
    st = ao.init_state(cbs, hyper, D)
    ll0 = rs.normal(-2.5, 1.0, n)
    lm0 = rs.normal(-3.5, 1.0, n)
    ll0[:10] = np.log(6e7 / 39.0)
    ll0[10:20] = 69.95                         # proposals cross +70: clipped by the exact path
    lm0[20:24] = 4.999                         # proposals cross log mu = 5: never accepted
    lm0[24:28] = 5.001                         # current target -inf: the first admissible proposal is accepted
    lm0[28:30] = 6.0
    ll0[30:34] = -69.9
    if case == "regular_sigma":
        Sigma = np.array([[1.3, 0.2, 0.1], [0.2, 2.1, -0.3], [0.1, -0.3, 0.8]])
    else:
        Sigma = np.array([[2e-13, 1e-14, 0.0], [1e-14, 3e-13, 0.0], [0.0, 0.0, 0.7]])
    beta = np.array([[-2.5, -3.5, 3.5], [0.1, -0.2, 0.05]])
    st["lam"], st["mu"] = np.exp(ll0), np.exp(lm0)
    st["beta"], st["Sigma"] = beta.copy(), Sigma.copy()
    src = PhiloxStreams(seed, 0, np.arange(n), S, D, 2, level1_variates=_device_fast_variates(seed, n, S))
    monkeypatch.setenv("CLV_SWEEP_CPT", cpt)
    with Sampler(x, t_x, T, X, log_s, model_dim=D, chains=1, n_mh_steps=S, seed=seed, rng="fast", sweep_mode="stream") as s:
        s.set_state(0, log_lambda=np.log(st["lam"]), log_mu=np.log(st["mu"]), log_eta=np.zeros(n), beta=beta, Sigma=Sigma)
        for step in range(1, 4):
            src.begin_sweep(step)
            with np.errstate(all="ignore"):
                ao.sweep(cbs, st, hyper, src, D, S)
            s.advance(1)
            dev = s.get_state(0)
            np.testing.assert_array_equal(dev["z"], st["z"].astype(float), err_msg=f"z, sweep {step}")
            np.testing.assert_allclose(dev["log_lambda"], np.log(st["lam"]), rtol=RTOL, atol=1e-9, err_msg=f"sweep {step}")
            np.testing.assert_allclose(dev["log_mu"], np.log(st["mu"]), rtol=RTOL, atol=1e-9, err_msg=f"sweep {step}")
            np.testing.assert_allclose(dev["tau"], st["tau"], rtol=RTOL, err_msg=f"sweep {step}")
            np.testing.assert_allclose(dev["beta"], st["beta"], rtol=RTOL, atol=1e-9)
            np.testing.assert_allclose(dev["Sigma"], st["Sigma"], rtol=RTOL, atol=1e-12)
    if case == "regular_sigma":
        assert np.log(st["lam"])[10:20].max() < 69.0 and (np.log(st["mu"])[24:30] <= 5.0).all()   # the chain did leave those corners


@pytest.mark.parametrize("n", [1, 2, 33, 127, 128, 129, 255, 256, 257, 641])
def test_tiny_and_ragged_customer_counts_vs_oracle(cdnow_full, monkeypatch, n):
    """Customer counts around the tile edges (128 customers per tile in the one-customer kernel, 256 in the
    two-customer one) down to a single customer: every sweep path -- k_sweep, k_sweep2 (forced with CLV_SWEEP_CPT)
    and the cooperative kernel -- against the oracle's replay of the strict Philox variates, bivariate and
    trivariate.  The rows are the first n customers WITH repeat purchases followed by their neighbours, so that
    the reference's initial state (bi:367-374: mean(x) / mean(t_x)) exists at n = 1."""
    d = cdnow_full
    order = np.argsort(d["x"] == 0, kind="stable")[:n]            # x > 0 first, input order otherwise
    if n > 2:
        order = np.sort(np.concatenate([order[: n // 2], np.flatnonzero(d["x"] == 0)[: n - n // 2]]))
    X = np.column_stack([np.ones(n), d["first_sales_scaled"][order]])
    cbs = ao.Cbs(x=d["x"][order].astype(np.int64), t_x=d["t_x"][order], T_cal=d["T_cal"][order], X=X, log_s=d["log_s"][order])
    seed, S = 31, 20
    for D in ((2,) if n == 1 else (2, 3)):                        # var(log_s) of one customer is undefined (tri:494)
        ora = ao.run_chain(cbs, ao.default_hyper(2, D), PhiloxStreams(seed, 0, np.arange(n), S, D, 2), mcmc=3, burnin=1,
                           thin=1, D=D, n_mh_steps=S)
        for cpt, mode in (("1", "stream"), ("2", "stream"), ("1", "persistent")):
            monkeypatch.setenv("CLV_SWEEP_CPT", cpt)
            with Sampler(cbs.x, cbs.t_x, cbs.T_cal, X, cbs.log_s if D == 3 else None, model_dim=D, chains=1, n_mh_steps=S,
                         seed=seed, rng="strict", sweep_mode=mode) as s:
                out = s.run(1, 3, 1)
            msg = f"n={n} D={D} cpt={cpt} {mode}"
            np.testing.assert_array_equal(out["level_1"][0][:, :, 3], ora["level_1"][:, :, 3], err_msg=msg)
            np.testing.assert_allclose(out["level_1"][0], ora["level_1"], rtol=RTOL, err_msg=msg)
            np.testing.assert_allclose(out["level_2"][0], ora["level_2"], rtol=RTOL, atol=1e-9, err_msg=msg)


@pytest.mark.parametrize("D,cov", [(2, []), (2, ["first_sales_scaled", "age_scaled"]), (3, ["gender_F", "age_scaled", "first_sales_scaled"]),
                                   (2, ["first_sales_scaled", "age_scaled", "gender_F", "first_sales_scaled", "age_scaled", "gender_F"])])
def test_column_wise_data_entry_equals_the_matrix_entry(cdnow_full, D, cov):
    """clv_set_data_columns (the covariate columns as the DataFrame holds them, intercept implicit -- what the drop-in
    entry points use) against clv_set_data (the (N, K) design matrix of bi:468-470): same initialisation statistics and
    the same chains, bit for bit; K = 1 (no covariate) to K = 7 (the general initialisation kernel)."""
    d = cdnow_full
    n = d["x"].size
    cols = [d[c].astype(float) for c in cov]
    X = np.column_stack([np.ones(n)] + cols)
    log_s = d["log_s"] if D == 3 else None
    outs = []
    for design in (X, cols):
        with Sampler(d["x"], d["t_x"], d["T_cal"], design, log_s, model_dim=D, chains=2, seed=11) as s:
            outs.append((s.init_stats, s.run(2, 3, 1)))
    for k in ("lam_init", "mean_mu_init", "mean_log_s", "omega2", "max_abs_x"):
        assert outs[0][0][k] == outs[1][0][k], k
    np.testing.assert_array_equal(outs[0][0]["xtx"], outs[1][0]["xtx"])
    for k in ("level_1", "level_2", "loglik_sum"):
        np.testing.assert_array_equal(outs[0][1][k], outs[1][1][k], err_msg=k)


def test_caller_provided_level1_array(cdnow_abe):
    """Sampler.run(out=...): the level-1 draws land in the caller's array (reusable over runs); same values as the
    array the run allocates itself; a wrong shape / dtype / layout is refused before any device work."""
    d = cdnow_abe
    n = 1200
    args = (d["x"][:n], d["t_x"][:n], d["T_cal"][:n], [d["first_sales_scaled"][:n].astype(float)])
    with Sampler(*args, chains=2, seed=8) as s:
        ref = s.run(3, 6, 2)
    buf = np.full((2, 3, n, 4), np.nan)
    with Sampler(*args, chains=2, seed=8) as s:
        out = s.run(3, 6, 2, out=buf)
        assert out["level_1"] is buf
        for bad in (np.empty((2, 3, n, 5)), np.empty((2, 3, n, 4), dtype=np.float32), np.empty((2, 3, 4, n)).transpose(0, 1, 3, 2)):
            with pytest.raises(ValueError, match="out must be"):
                s.run(3, 6, 2, out=bad)
    np.testing.assert_array_equal(buf, ref["level_1"])
    np.testing.assert_array_equal(out["level_2"], ref["level_2"])
