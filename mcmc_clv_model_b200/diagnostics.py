"""Convergence diagnostics on level-2 draws (host NumPy; replaces the `az.summary` / `az.plot_autocorr` calls of
bivariate/analysis_abe.py:651-706 -- arviz is not a dependency here).

ess_bulk / ess_tail / rhat follow Vehtari, Gelman, Simpson, Carpenter & Buerkner (2021), i.e. what arviz computes:
split chains, rank-normalisation (bulk, R-hat), Geyer's initial monotone positive sequence on the combined
autocovariance; ess_tail = min over the 5 % / 95 % quantile indicators.  ess_geyer is the per-chain
initial-positive-sequence estimate, summed over chains (the "crude Geyer" figure BASELINE.md quotes for the
reference).  autocorr is the per-chain autocorrelation function `az.plot_autocorr` draws (lags 0..max_lag).
"""
from __future__ import annotations

import numpy as np
from scipy import stats


def _autocov(x):
    """Autocovariance of each row of x (chains, n) by FFT, biased (divide by n)."""
    n = x.shape[1]
    m = 1 << (2 * n - 1).bit_length()
    xc = x - x.mean(axis=1, keepdims=True)
    f = np.fft.rfft(xc, m, axis=1)
    ac = np.fft.irfft(f * np.conj(f), m, axis=1)[:, :n]
    return ac / n


def _split(x):
    n = x.shape[1] // 2
    return np.concatenate([x[:, :n], x[:, -n:]], axis=0)


def _rank_normalise(x):
    r = stats.rankdata(x.ravel(), method="average").reshape(x.shape)
    return stats.norm.ppf((r - 0.375) / (x.size + 0.25))


def _ess_core(x):
    """Multi-chain ESS of x (chains, n) (Stan's algorithm)."""
    m, n = x.shape
    if n < 4:
        return float("nan")
    acov = _autocov(x)
    chain_var = acov[:, 0] * n / (n - 1.0)
    W = chain_var.mean()
    var_plus = W * (n - 1.0) / n
    if m > 1:
        var_plus += x.mean(axis=1).var(ddof=1)
    if not var_plus > 0:
        return float("nan")
    rho = 1.0 - (W - acov.mean(axis=0)) / var_plus
    rho[0] = 1.0
    # Geyer: sums of adjacent pairs, truncated at the first negative pair, made monotone
    T = (n - 1) // 2
    pairs = rho[0:2 * T:2] + rho[1:2 * T + 1:2]
    neg = np.flatnonzero(pairs < 0)
    k = neg[0] if neg.size else pairs.size
    pairs = np.minimum.accumulate(pairs[:k])
    tau = -1.0 + 2.0 * pairs.sum()
    tau = max(tau, 1.0 / np.log10(m * n))
    return m * n / tau


def ess_bulk(x):
    """x: (chains, n_draws) of one scalar parameter."""
    x = np.asarray(x, dtype=float)
    return _ess_core(_rank_normalise(_split(x)))


def ess_geyer(x):
    """Sum over chains of the single-chain initial-positive-sequence ESS."""
    x = np.asarray(x, dtype=float)
    return float(sum(_ess_core(x[c:c + 1]) for c in range(x.shape[0])))


def ess_mean(x):
    """Split-chain ESS of the mean, no rank normalisation (the denominator of arviz's mcse_mean)."""
    return _ess_core(_split(np.asarray(x, dtype=float)))


def ess_quantile(x, prob):
    """ESS of the indicator x <= quantile(x, prob) (split chains), as arviz's ess(method="quantile")."""
    x = np.asarray(x, dtype=float)
    return _ess_core(_split((x <= np.quantile(x, prob)).astype(float)))


def ess_tail(x):
    """min of the 5 % and 95 % quantile ESS (arviz `ess_tail`)."""
    return float(min(ess_quantile(x, 0.05), ess_quantile(x, 0.95)))


def autocorr(x, max_lag=100):
    """Autocorrelation function of each chain: x (chains, n) -> (chains, max_lag + 1), rho_0 = 1
    (what `az.plot_autocorr` draws, bivariate/analysis_abe.py:694-706)."""
    x = np.atleast_2d(np.asarray(x, dtype=float))
    ac = _autocov(x)
    k = min(max_lag + 1, x.shape[1])
    with np.errstate(invalid="ignore", divide="ignore"):
        return ac[:, :k] / ac[:, :1]


def rhat(x):
    """Rank-normalised split R-hat; the maximum of the bulk and the folded (tail) version, as arviz."""
    x = np.asarray(x, dtype=float)

    def _r(v):
        v = _rank_normalise(_split(v))
        n = v.shape[1]
        W = v.var(axis=1, ddof=1).mean()
        B = n * v.mean(axis=1).var(ddof=1)
        return float(np.sqrt(((n - 1.0) / n * W + B / n) / W))
    return max(_r(x), _r(np.abs(x - np.median(x))))


def hdi(flat, prob=0.94):
    """Narrowest interval holding `prob` of the draws (arviz's default hdi_prob = 0.94: hdi_3%, hdi_97%)."""
    v = np.sort(np.asarray(flat, dtype=float).ravel())
    n = v.size
    k = int(np.floor(prob * n))
    if k < 1 or k >= n:
        return float(v[0]), float(v[-1])
    w = v[k:] - v[:n - k]
    i = int(np.argmin(w))
    return float(v[i]), float(v[i + k])


def summarize(level_2, names=None):
    """level_2: list/array (chains, n_draws, P) -> dict per column with az.summary's columns (mean, sd, hdi_3%, hdi_97%,
    mcse_mean, ess_bulk, ess_tail, r_hat) plus q2.5/q50/q97.5 and the crude Geyer ESS."""
    a = np.asarray(level_2, dtype=float)
    out = {}
    for j in range(a.shape[2]):
        col = a[:, :, j]
        flat = col.ravel()
        eb, eg = ess_bulk(col), ess_geyer(col)
        q = np.percentile(flat, [2.5, 50, 97.5])
        lo, hi = hdi(flat)
        sd = float(flat.std(ddof=1))
        out[names[j] if names else j] = dict(mean=float(flat.mean()), sd=sd, q025=float(q[0]),
                                             q50=float(q[1]), q975=float(q[2]), hdi_3=lo, hdi_97=hi, ess_bulk=float(eb),
                                             ess_tail=ess_tail(col), ess_geyer=float(eg),
                                             mcse_mean=float(sd / np.sqrt(max(eg, 1.0))),
                                             mcse_mean_split=float(sd / np.sqrt(max(ess_mean(col), 1.0))),
                                             rhat=rhat(col), r_hat=rhat(col))
    return out


def summary_table(level_2, names=None, round_to=4):
    """The table `az.summary(idata, var_names=["level_2"], round_to=4)` prints (bivariate/analysis_abe.py:687-690) as a
    pandas DataFrame: index = parameter, columns mean, sd, hdi_3%, hdi_97%, mcse_mean, ess_bulk, ess_tail, r_hat."""
    import pandas as pd
    s = summarize(level_2, names)
    rows = {k: {"mean": v["mean"], "sd": v["sd"], "hdi_3%": v["hdi_3"], "hdi_97%": v["hdi_97"],
                "mcse_mean": v["mcse_mean_split"], "ess_bulk": v["ess_bulk"], "ess_tail": v["ess_tail"], "r_hat": v["r_hat"]}
            for k, v in s.items()}
    return pd.DataFrame.from_dict(rows, orient="index").round(round_to)


def min_ess(level_2, kind="bulk"):
    s = summarize(level_2)
    key = {"bulk": "ess_bulk", "tail": "ess_tail"}.get(kind, "ess_geyer")
    return min(v[key] for v in s.values())
