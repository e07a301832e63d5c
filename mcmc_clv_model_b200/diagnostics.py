"""Convergence diagnostics on level-2 draws (host NumPy; replaces the `az.summary` call of
bivariate/analysis_abe.py:651-693 -- arviz is not a dependency here).

ess_bulk / rhat follow Vehtari, Gelman, Simpson, Carpenter & Buerkner (2021): rank-normalised,
split chains, Geyer's initial monotone positive sequence on the combined autocovariance.
ess_geyer is the per-chain initial-positive-sequence estimate, summed over chains (the "crude Geyer"
figure BASELINE.md quotes for the reference).
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


def rhat(x):
    x = _rank_normalise(_split(np.asarray(x, dtype=float)))
    m, n = x.shape
    W = x.var(axis=1, ddof=1).mean()
    B = n * x.mean(axis=1).var(ddof=1)
    return float(np.sqrt(((n - 1.0) / n * W + B / n) / W))


def summarize(level_2, names=None):
    """level_2: list/array (chains, n_draws, P) -> dict per column: mean, sd, q2.5, q50, q97.5, ess_bulk,
    ess_geyer, mcse_mean, rhat."""
    a = np.asarray(level_2, dtype=float)
    out = {}
    for j in range(a.shape[2]):
        col = a[:, :, j]
        flat = col.ravel()
        eb, eg = ess_bulk(col), ess_geyer(col)
        q = np.percentile(flat, [2.5, 50, 97.5])
        out[names[j] if names else j] = dict(mean=float(flat.mean()), sd=float(flat.std(ddof=1)), q025=float(q[0]),
                                             q50=float(q[1]), q975=float(q[2]), ess_bulk=float(eb), ess_geyer=float(eg),
                                             mcse_mean=float(flat.std(ddof=1) / np.sqrt(max(eg, 1.0))),
                                             rhat=rhat(col))
    return out


def min_ess(level_2, kind="bulk"):
    s = summarize(level_2)
    key = "ess_bulk" if kind == "bulk" else "ess_geyer"
    return min(v[key] for v in s.values())
