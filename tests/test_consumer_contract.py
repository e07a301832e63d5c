"""Consumer contract: the reference's own analysis helpers (pure NumPy, src/models/utils/analysis_bi_helpers.py) must run
unchanged on the dict our drop-in `mcmc_draw_parameters` returns.  tests/golden/gpu_draws_small.pkl was produced on a B200
by tools/make_gpu_draws_fixture.py; the helpers are imported from /root/reference when it exists (build container)."""
import os
import pickle
import sys

import numpy as np
import pytest

from conftest import GOLDEN, load_golden

REF = "/root/reference"


@pytest.fixture(scope="module")
def fixture():
    with open(os.path.join(GOLDEN, "gpu_draws_small.pkl"), "rb") as f:
        return pickle.load(f)


def test_layout_of_the_pickled_dict(fixture):
    draws, n = fixture["draws"], fixture["n"]
    assert set(draws) == {"level_1", "level_2", "log_likelihood"}                       # bi:503-504
    assert isinstance(draws["level_1"], list) and len(draws["level_1"]) == 2
    for l1, l2 in zip(draws["level_1"], draws["level_2"]):
        assert l1.shape == (30, n, 4) and l1.dtype == np.float64 and l1.flags.c_contiguous  # (n_draws, N, 4) = lambda, mu, tau, z
        assert l2.shape == (30, 2 * 2 + 3)                                               # beta.T.ravel(), S00, S01, S11
        assert set(np.unique(l1[:, :, 3])) <= {0.0, 1.0} and np.all(l1[:, :, :2] > 0)
    assert np.array(draws["level_2"]).shape == (2, 30, 7)                               # analysis_abe.py:651-675
    assert isinstance(draws["log_likelihood"], np.floating)
    assert fixture["x_star"].shape == (60, n) and fixture["x_star"].dtype == np.int64    # bi:546


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_reference_analysis_helpers_run_on_our_draws(fixture):
    sys.path.insert(0, REF)
    try:
        import importlib
        for m in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[m]                       # the repo's own `src` package shadows the reference's
        h = importlib.import_module("src.models.utils.analysis_bi_helpers")
        assert h.__file__.startswith(REF)
        import pandas as pd
        draws, n = fixture["draws"], fixture["n"]
        d = load_golden("cdnow_abe.npz")
        cbs = pd.DataFrame({k: d[k][:n] for k in ("x", "t_x", "T_cal", "x_star")})
        names = ["b_l0", "b_l1", "b_m0", "b_m1", "var_l", "cov", "var_m"]
        s = h.summarize_level2(draws["level_2"][0], names)                               # analysis_abe.py:146-147
        assert s.shape == (7, 3) and np.isfinite(s.to_numpy()).all()
        lam, mu = h.post_mean_lambdas(draws), h.post_mean_mus(draws)
        assert lam.shape == (n,) and mu.shape == (n,) and (lam > 0).all()
        corr = h.extract_correlation(np.vstack(draws["level_2"]))
        assert np.all(np.abs(corr) <= 1)
        ll = h.chain_total_loglik(draws["level_1"], cbs)
        assert np.isfinite(ll) and ll < 0
        t4 = h.compute_table4(draws, fixture["x_star"])
        assert "P(alive at T_cal)" in t4.columns and len(t4) == 24
        # our device summaries are the same numbers the helper computes on the host
        allv = np.concatenate(draws["level_1"], axis=0)
        np.testing.assert_allclose(lam, allv[:, :, 0].mean(axis=0))
    finally:
        sys.path.remove(REF)
        for m in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[m]
