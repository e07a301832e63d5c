"""NumPy restatement of the device RNG contract (TEST INFRASTRUCTURE ONLY).

This is *not* a restatement of reference code: the reference draws from NumPy's
PCG64 (`/root/reference/src/models/bivariate/mcmc.py:486`).  It restates the
counter-based Philox4x32-10 contract of ``mcmc_clv_model_b200/csrc/clv_rng.cuh``
so that the GPU's "strict f64" Philox mode can be replayed through the oracle
(and through the reference's own block functions) variate for variate.

Contract
--------
key     = (lo32(seed), hi32(seed))
counter = (customer_gid, sweep, slot, domain | chain << 4)
domain  : 0 sampler level-1, 1 level-2 draw, 2 forecast, 3 synthetic generator
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

DOM_SAMPLER, DOM_LEVEL2, DOM_FORECAST, DOM_GENERATOR = 0, 1, 2, 3

TWO_PI = 6.283185307179586476925286766559


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs broadcastable, treated as uint32.
    Returns four uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(
        *(np.asarray(v, dtype=np.uint64) & MASK32 for v in (c0, c1, c2, c3))
    )
    c0, c1, c2, c3 = (c.copy() for c in (c0, c1, c2, c3))
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def chain_key(seed: int, chain: int = 0):
    """The key is the seed; the chain index travels in the counter (dom_word)."""
    s = int(seed) & 0xFFFFFFFFFFFFFFFF
    return s & 0xFFFFFFFF, s >> 32


def dom_word(domain: int, chain: int = 0) -> int:
    return (domain | (int(chain) << 4)) & 0xFFFFFFFF


def u53(a, b):
    """53-bit uniform strictly inside (0,1) from two uint32 words."""
    a = np.asarray(a, dtype=np.uint64)
    b = np.asarray(b, dtype=np.uint64)
    m = (a >> np.uint64(5)) * np.uint64(1 << 26) + (b >> np.uint64(6))  # 53 bits
    return (m.astype(np.float64) + 0.5) * (2.0 ** -53)


def u32(a):
    """32-bit uniform strictly inside (0,1) (strict-f64 mode)."""
    return (np.asarray(a, dtype=np.float64) + 0.5) * (2.0 ** -32)


def u24(a):
    """24-bit uniform strictly inside (0,1) from the TOP 24 bits of a word (the level-1 proposal transforms)."""
    return ((np.asarray(a, dtype=np.uint64) >> np.uint64(8)).astype(np.float64) + 0.5) * (2.0 ** -24)


def low_bytes(w0, w1, w2, w3):
    """The accept word of a Metropolis step: byte 0 of the first four words of the step -> bytes 0..3."""
    b = [np.asarray(w, dtype=np.uint64) & np.uint64(0xFF) for w in (w0, w1, w2, w3)]
    return b[0] | (b[1] << np.uint64(8)) | (b[2] << np.uint64(16)) | (b[3] << np.uint64(24))


def step_words(seed, chain, gids, sweep, step):
    """The four words Metropolis step `step` consumes (clv_rng.cuh, "word layout of the Metropolis steps"): the Philox
    block of slot 1 + step; (x, y) = (V, angle) for the log-lambda proposal, (z, w) for log mu."""
    k0, k1 = chain_key(seed, chain)
    c3 = dom_word(DOM_SAMPLER, chain)
    return philox4x32_10(gids, sweep, 1 + step, c3, k0, k1)


def t3_from_words(rv, rb):
    """Student-t(3) without rejection from two uniforms: with a Box-Muller pair R (cos th, sin th) and an independent
    chi-square(2) = 2 E3,  t / sqrt(3) = cos th / sqrt(sin^2 th + E3 / E1), and the ratio of two independent Exp(1)
    variables is V / (1 - V) with V uniform."""
    v, ang = u24(rv), TWO_PI * u24(rb)
    return np.sqrt(3.0) * np.cos(ang) / np.sqrt(np.sin(ang) ** 2 + v / (1.0 - v))


def normal_pair_u53(r0, r1, r2, r3):
    ua, ub = u53(r0, r1), u53(r2, r3)
    r = np.sqrt(-2.0 * np.log(ua))
    ang = TWO_PI * ub
    return r * np.cos(ang), r * np.sin(ang)


# ---------------------------------------------------------------------------
# sampler domain
# ---------------------------------------------------------------------------
def sampler_variates(seed, chain, gids, sweep, n_mh_steps, with_eta=False):
    """All per-customer variates of one sweep in strict-f64 mode.

    Returns dict: u_z, e_tau, u_tau (N,), t3_l, t3_m, u_acc (S,N), [n_eta (N,)].
    e_tau and u_tau derive from the same 53-bit uniform (only one is consumed,
    depending on z)."""
    k0, k1 = chain_key(seed, chain)
    gids = np.asarray(gids, dtype=np.uint64)
    c3 = dom_word(DOM_SAMPLER, chain)
    r = philox4x32_10(gids, sweep, 0, c3, k0, k1)
    u_z = u53(r[0], r[1])
    u_t = u53(r[2], r[3])
    out = dict(u_z=u_z, u_tau=u_t, e_tau=-np.log(u_t))
    S = n_mh_steps
    t3_l = np.empty((S, gids.size))
    t3_m = np.empty((S, gids.size))
    u_acc = np.empty((S, gids.size))
    for s in range(S):
        w = step_words(seed, chain, gids, sweep, s)
        t3_l[s] = t3_from_words(w[0], w[1])
        t3_m[s] = t3_from_words(w[2], w[3])
        u_acc[s] = u32(low_bytes(w[0], w[1], w[2], w[3]))
    out.update(t3_l=t3_l, t3_m=t3_m, u_acc=u_acc)
    if with_eta:
        e = philox4x32_10(gids, sweep, 1 + 2 * S, c3, k0, k1)
        out["n_eta"] = normal_pair_u53(e[0], e[1], e[2], e[3])[0]
    return out


# ---------------------------------------------------------------------------
# level-2 domain
# ---------------------------------------------------------------------------
def level2_normal(seed, chain, sweep, idx):
    k0, k1 = chain_key(seed, chain)
    r = philox4x32_10(np.asarray([idx]), sweep, 0, dom_word(DOM_LEVEL2, chain), k0, k1)
    return float(normal_pair_u53(r[0], r[1], r[2], r[3])[0][0])


def level2_chi2(seed, chain, sweep, idx, df):
    """chi2(df) = 2*Gamma(df/2) by Marsaglia-Tsang (df >= 2), f64."""
    k0, k1 = chain_key(seed, chain)
    a = 0.5 * float(df)
    d = a - 1.0 / 3.0
    c = 1.0 / np.sqrt(9.0 * d)
    attempt = 0
    while True:
        ra = philox4x32_10(np.asarray([idx]), sweep, 2 * attempt, dom_word(DOM_LEVEL2, chain), k0, k1)
        rb = philox4x32_10(np.asarray([idx]), sweep, 2 * attempt + 1, dom_word(DOM_LEVEL2, chain), k0, k1)
        x = float(normal_pair_u53(ra[0], ra[1], ra[2], ra[3])[0][0])
        u = float(u53(rb[0], rb[1])[0])
        attempt += 1
        v = 1.0 + c * x
        if v <= 0.0:
            continue
        v = v * v * v
        if np.log(u) < 0.5 * x * x + d - d * v + d * np.log(v):
            return 2.0 * d * v


def level2_variates(seed, chain, sweep, D, K, nu_n):
    """Variates of one level-2 draw: Bartlett normals, chi2, beta normals.
    Index plan: [0, n_tril) normals; [16, 16+D) chi2; [32, 32+D*K) beta normals."""
    n_tril = D * (D - 1) // 2
    iw_norm = np.array([level2_normal(seed, chain, sweep, j) for j in range(n_tril)])
    iw_chi2 = np.array(
        [level2_chi2(seed, chain, sweep, 16 + i, nu_n - D + 1 + i) for i in range(D)]
    )
    beta_norm = np.array([level2_normal(seed, chain, sweep, 32 + j) for j in range(D * K)])
    return dict(iw_norm=iw_norm, iw_chi2=iw_chi2, beta_norm=beta_norm)


# ---------------------------------------------------------------------------
# forecast domain
# ---------------------------------------------------------------------------
def forecast_uniform(seed, gids, draw):
    """One block serves the two draws 2g, 2g+1 (global, chain-major draw index): words (0,1) / (2,3)."""
    k0, k1 = chain_key(seed, 0)
    r = philox4x32_10(np.asarray(gids, dtype=np.uint64), draw >> 1, 0, DOM_FORECAST, k0, k1)
    return u53(r[2], r[3]) if draw & 1 else u53(r[0], r[1])


PTRS_MIN_MEAN = 60.0
# forecast-domain slot ranges (clv_forecast.cuh): one counter -> one variate, ranges disjoint
FC_SLOT_PTRS, FC_SLOT_WEEK, FC_SLOT_WEEK_PTRS, FC_SLOT_SPEND = 0x00010000, 0x00020000, 0x01000000, 0x80000000


def forecast_poisson_ptrs(seed, gid, draw, lam):
    """Device contract for means >= PTRS_MIN_MEAN: Hoermann's PTRS (the algorithm NumPy's Generator.poisson uses for
    lam >= 10; numpy/random/src/distributions/distributions.c, third-party, not in the reference tree); attempt t draws
    its two uniforms from Philox block (gid, draw, FC_SLOT_PTRS + t, DOM_FORECAST)."""
    from math import lgamma, log, sqrt, floor, fabs
    k0, k1 = chain_key(seed, 0)
    slam, loglam = sqrt(lam), log(lam)
    b = 0.931 + 2.53 * slam
    a = -0.059 + 0.02483 * b
    invalpha = 1.1239 + 1.1328 / (b - 3.4)
    vr = 0.9277 - 3.6224 / (b - 2.0)
    t = 0
    while True:
        r = philox4x32_10(np.asarray([gid], dtype=np.uint64), draw, FC_SLOT_PTRS + t, DOM_FORECAST, k0, k1)
        t += 1
        U = float(u53(r[0], r[1])[0]) - 0.5
        V = float(u53(r[2], r[3])[0])
        us = 0.5 - fabs(U)
        kd = floor((2.0 * a / us + b) * U + lam + 0.43)
        if us >= 0.07 and V <= vr:
            return int(kd)
        if kd < 0 or (us < 0.013 and V > us):
            continue
        if log(V) + log(invalpha) - log(a / (us * us) + b) <= -lam + kd * loglam - lgamma(kd + 1.0):
            return int(kd)


def forecast_spend_normal(seed, gid, draw, j):
    """j-th per-transaction normal of cell (draw, gid)."""
    k0, k1 = chain_key(seed, 0)
    r = philox4x32_10(np.asarray([gid], dtype=np.uint64), draw, FC_SLOT_SPEND + j // 2, DOM_FORECAST, k0, k1)
    c, s = normal_pair_u53(r[0], r[1], r[2], r[3])
    return float(c[0] if j % 2 == 0 else s[0])


def weekly_uniform(seed, gids, draw, w):
    """Uniform of week index w for (customer, global draw): block (gid, draw, FC_SLOT_WEEK + w//4), word w%4, 32 bits."""
    k0, k1 = chain_key(seed, 0)
    r = philox4x32_10(np.asarray(gids, dtype=np.uint64), draw, FC_SLOT_WEEK + w // 4, DOM_FORECAST, k0, k1)
    return u32(r[w % 4])
