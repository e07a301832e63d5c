"""Host-side arithmetic that must not depend on how customers are sharded.

The reference initialises every chain from four data means (bi:367-374, tri:488-499).  Summing doubles
in shard order would make those means, and through the sampler's sensitivity every later draw, depend
on the GPU count.  Here each term is rounded once to a fixed-point integer (>= 41 bits below the largest
magnitude) and the integers are added exactly (int64 tiles, Python ints above), so any partition of the
customers gives bit-identical statistics.  Partials are plain Python ints, all-reduced by the caller.
"""
from __future__ import annotations

import math
from fractions import Fraction

import numpy as np

_TILE = 1024


def fx_bits(max_abs: float) -> int:
    """Fixed-point fraction bits such that a 1024-term tile of |v| <= max_abs fits in int64."""
    if not max_abs > 1.0:
        return 51
    mant, ex = math.frexp(max_abs)          # max_abs = mant * 2**ex, mant in [0.5, 1)
    return 51 - (ex - 1 if mant == 0.5 else ex)   # 51 - ceil(log2(max_abs)), exactly


def exact_partial(v: np.ndarray, bits: int) -> int:
    """sum(round(v * 2**bits)) as an exact Python int."""
    v = np.ascontiguousarray(v, dtype=np.float64).ravel()
    q = np.rint(np.ldexp(v, bits)).astype(np.int64)
    pad = (-q.size) % _TILE
    if pad:
        q = np.concatenate([q, np.zeros(pad, dtype=np.int64)])
    tiles = q.reshape(-1, _TILE).sum(axis=1)
    return int(np.sum(tiles.astype(object))) if tiles.size else 0


def from_fx(total: int, bits: int) -> float:
    """Correctly rounded total * 2**-bits."""
    return float(Fraction(total, 1 << bits)) if bits >= 0 else float(total * (1 << -bits))


class ExactSum:
    """Sum of doubles independent of partitioning.  `allreduce_max(float)->float` and
    `allreduce_sum_int(int)->int` are identities when unsharded."""

    def __init__(self, allreduce_max=None, allreduce_sum_int=None):
        self.amax = allreduce_max or (lambda x: x)
        self.asum = allreduce_sum_int or (lambda x: x)

    def __call__(self, v: np.ndarray) -> float:
        v = np.asarray(v, dtype=np.float64)
        m = float(np.max(np.abs(v))) if v.size else 0.0
        if not math.isfinite(m):
            raise ValueError("non-finite value in an initialisation statistic")
        bits = fx_bits(self.amax(m))
        return from_fx(self.asum(exact_partial(v, bits)), bits)


def init_statistics(x, t_x, T_cal, X, log_s, n_global, esum: ExactSum | None = None):
    """Global initialisation statistics of bi:367-374 / tri:488-499 from one shard's columns.
    Returns dict(lam_init, mean_mu_init, mean_log_s, omega2, max_abs_x, xtx)."""
    esum = esum or ExactSum()
    x = np.asarray(x, dtype=np.float64)
    t_x = np.asarray(t_x, dtype=np.float64)
    T_cal = np.asarray(T_cal, dtype=np.float64)
    n = float(n_global)
    mean_x = esum(x) / n
    mean_t = esum(np.where(t_x == 0, T_cal, t_x)) / n                      # bi:368
    lam_init = mean_x / mean_t
    mean_mu = esum(1.0 / (t_x + 0.5 / lam_init)) / n                        # bi:370, 374
    K = X.shape[1]
    xtx = np.empty((K, K))
    for a in range(K):
        for b in range(a, K):
            xtx[a, b] = xtx[b, a] = esum(X[:, a] * X[:, b])                 # bi:248
    out = dict(lam_init=lam_init, mean_mu_init=mean_mu, mean_log_s=0.0, omega2=1.0,
               max_abs_x=math.sqrt(max(1.0, esum.amax(float(np.max(X * X)) if X.size else 1.0))), xtx=xtx)
    if log_s is not None:
        log_s = np.asarray(log_s, dtype=np.float64)
        m = esum(log_s) / n                                                 # tri:499
        out["mean_log_s"] = m
        out["omega2"] = esum((log_s - m) ** 2) / (n - 1.0)                  # tri:494 (pandas var, ddof=1)
    return out
