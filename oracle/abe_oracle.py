"""NumPy restatement of the reference sampler's algorithm, block by block.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Every function cites the
reference lines it follows (paths relative to ``/root/reference``; "bi" =
``src/models/bivariate/mcmc.py``, "tri" = ``src/models/trivariate/mcmc.py``).

Unlike the reference, no function here owns a random generator: all variates
come from a *stream source* (``oracle/streams.py``) so that the same arithmetic
can be driven by (a) NumPy's PCG64 in the reference's call order (reproduces the
reference's chains bit for bit), (b) injected arrays, (c) the device's Philox
contract.  Data are plain arrays, not DataFrames.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import numpy as np


@dataclass
class Cbs:
    """Calibration data of one customer shard (bi:55-69, bi:467-470)."""
    x: np.ndarray          # (N,) int64 repeat transactions
    t_x: np.ndarray        # (N,) f64 recency
    T_cal: np.ndarray      # (N,) f64 calibration length
    X: np.ndarray          # (N,K) f64 design matrix, column 0 == 1
    log_s: Optional[np.ndarray] = None   # (N,) f64, trivariate only (tri:329)

    @property
    def N(self):
        return self.x.shape[0]

    @property
    def K(self):
        return self.X.shape[1]


def default_hyper(K: int, D: int):
    """Diffuse NIW prior (bi:474-479 for D=2; tri:622-626 for D=3)."""
    nu0 = (3 + K) if D == 2 else (4 + K)
    return dict(beta_0=np.zeros((K, D)), A_0=np.eye(K) * 0.01, nu_00=nu0,
                gamma_00=nu0 * np.eye(D))


# ---------------------------------------------------------------------------
# a1  alive indicator  (bi:193-200, tri:272-277)
# ---------------------------------------------------------------------------
def p_alive(t_x, T_cal, lam, mu):
    ml = mu + lam
    e = np.exp(-(ml * (T_cal - t_x)))
    return (ml * e) / (ml * e + mu * (1.0 - e))


def draw_z(t_x, T_cal, lam, mu, u):
    return u < p_alive(t_x, T_cal, lam, mu)


# ---------------------------------------------------------------------------
# a2  dropout time  (bi:203-227, tri:280-304)
# ---------------------------------------------------------------------------
def draw_tau(t_x, T_cal, lam, mu, z, e, u):
    """``e``/``u`` are full-length; only e[z] and u[~z] are consumed."""
    ml = mu + lam
    tau = np.empty_like(t_x, dtype=float)
    a = np.flatnonzero(z)
    if a.size:                                  # bi:215-217
        tau[a] = T_cal[a] + (1.0 / mu[a]) * e[a]
    c = np.flatnonzero(~z)
    if c.size:                                  # bi:220-226
        m = ml[c]
        m_tx = np.minimum(700.0, m * t_x[c])
        m_T = np.minimum(700.0, m * T_cal[c])
        uc = u[c]
        tau[c] = -np.log((1 - uc) * np.exp(-m_tx) + uc * np.exp(-m_T)) / m
    return tau


# ---------------------------------------------------------------------------
# a3  level-1 MH  (bi:268-339, tri:387-458)
# ---------------------------------------------------------------------------
def log_posterior(ll, lm, x, T_cal, z, tau, mean_l, mean_m, P):
    """bi:291-310.  P = inv(Sigma); only P[0,0], P[0,1], P[1,1] enter (Q4)."""
    dl = ll - mean_l
    dm = lm - mean_m
    lik = x * ll + (1 - z) * lm - (np.exp(ll) + np.exp(lm)) * (z * T_cal + (1 - z) * tau)
    prior = -0.5 * (dl ** 2 * P[0, 0] + 2 * dl * dm * P[0, 1] + dm ** 2 * P[1, 1])
    res = lik + prior
    return np.where(lm > 5.0, -np.inf, res)


def draw_level_1(cbs: Cbs, lam, mu, z, tau, beta, Sigma, src, n_mh_steps=20,
                 return_logs=False):
    P = np.linalg.inv(Sigma)                    # bi:283
    mean = cbs.X @ beta                         # bi:284
    ll = np.log(lam)
    lm = np.log(mu)
    N = ll.size
    zi = z  # bool array; (1 - z) works on bool as in the reference
    cur = log_posterior(ll, lm, cbs.x, cbs.T_cal, zi, tau, mean[:, 0], mean[:, 1], P)
    n_acc = 0
    with np.errstate(over="ignore", invalid="ignore"):
        for _ in range(n_mh_steps):             # bi:314-335
            pl = ll + Sigma[0, 0] * src.t3(N)
            pm = lm + Sigma[1, 1] * src.t3(N)
            pl = np.clip(pl, -70.0, 70.0)
            pm = np.clip(pm, -70.0, 70.0)
            prop = log_posterior(pl, pm, cbs.x, cbs.T_cal, zi, tau, mean[:, 0], mean[:, 1], P)
            acc = np.exp(prop - cur) > src.u_acc(N)
            ll[acc] = pl[acc]
            lm[acc] = pm[acc]
            cur[acc] = prop[acc]
            n_acc += int(acc.sum())
    if return_logs:
        return ll, lm, n_acc
    return np.exp(ll), np.exp(lm)               # bi:337-338


# ---------------------------------------------------------------------------
# a4  level-2 conjugate regression  (bi:233-262, tri:340-380)
# ---------------------------------------------------------------------------
def level_2_posterior(X, Y, hyper):
    A0, B0, nu0, S0 = hyper["A_0"], hyper["beta_0"], hyper["nu_00"], hyper["gamma_00"]
    V = np.linalg.inv(X.T @ X + A0)             # bi:248-249
    B_hat = V @ (X.T @ Y + A0 @ B0)             # bi:250
    E = Y - X @ B_hat                           # bi:253
    C = B_hat - B0
    S_n = S0 + E.T @ E + C.T @ A0 @ C           # bi:255
    return V, B_hat, S_n, nu0 + X.shape[0]      # bi:256


def inv_wishart_bartlett(S_n, normals, chi2):
    """SciPy 1.18.1 ``invwishart.rvs`` restated (third-party, not in the
    reference tree; ``scipy/stats/_multivariate.py:3637-3731``): lower-triangular
    A with standard normals below the diagonal in ``np.tril_indices(D,-1)``
    order and sqrt(chi2(nu-D+1+i)) on the diagonal; Sigma = (C A^-1)(C A^-1)^T
    with C = chol(S_n)."""
    D = S_n.shape[0]
    A = np.zeros((D, D))
    r, c = np.tril_indices(D, k=-1)
    A[r, c] = normals
    A[np.arange(D), np.arange(D)] = np.sqrt(chi2)
    C = np.linalg.cholesky(S_n)
    CA = np.linalg.solve(A.T, C.T).T            # C @ inv(A)
    return CA @ CA.T


def beta_noise_cholesky(Sigma, V, zvec):
    """chol(kron(Sigma, V)) @ z  ==  kron(chol Sigma, chol V) @ z."""
    L = np.kron(np.linalg.cholesky(Sigma), np.linalg.cholesky(V))
    return L @ zvec


def assemble_beta(B_hat, noise, compat="reference"):
    """bi:261 / tri:376-378.  ``noise`` is ordered d*K+k (kron(Sigma,V) order).
    compat="reference": added to ``B_hat.ravel()`` (ordered k*D+d) as the
    reference does.  compat="paper": the matrix-normal it meant."""
    K, D = B_hat.shape
    if compat == "reference":
        return (B_hat.ravel() + noise).reshape(K, D)
    return B_hat + noise.reshape(D, K).T


def draw_level_2(X, Y, hyper, src, compat="reference"):
    V, B_hat, S_n, nu_n = level_2_posterior(X, Y, hyper)
    Sigma = src.inv_wishart(nu_n, S_n)          # bi:258
    beta = src.beta(B_hat, Sigma, V, compat)    # bi:261
    return beta, Sigma


# ---------------------------------------------------------------------------
# a5  eta  (tri:306-333)
# ---------------------------------------------------------------------------
def draw_log_eta(log_s, X, beta, Sigma, omega2, n):
    prior_mean = (X @ beta)[:, 2]
    prior_var = Sigma[2, 2]
    post_var = 1.0 / (1.0 / omega2 + 1.0 / prior_var)
    post_mean = post_var * (log_s / omega2 + prior_mean / prior_var)
    return post_mean + np.sqrt(post_var) * n


# ---------------------------------------------------------------------------
# a6  chain driver  (bi:346-431, tri:465-574)
# ---------------------------------------------------------------------------
def init_state(cbs: Cbs, hyper, D):
    """bi:367-379 / tri:488-504.  Mutates hyper['beta_0'] row 0 like the reference."""
    lam_init = cbs.x.mean() / np.mean(np.where(cbs.t_x == 0, cbs.T_cal, cbs.t_x))
    lam = np.full(cbs.N, lam_init)
    mu = 1.0 / (cbs.t_x + 0.5 / lam_init)
    hyper["beta_0"][0, 0] = math.log(lam.mean())
    hyper["beta_0"][0, 1] = math.log(mu.mean())
    st = dict(lam=lam, mu=mu, beta=hyper["beta_0"].copy(), Sigma=hyper["gamma_00"].copy())
    if D == 3:
        st["eta"] = np.ones(cbs.N)
        n = cbs.N
        m = cbs.log_s.sum() / n
        st["omega2"] = float(((cbs.log_s - m) ** 2).sum() / (n - 1))   # pandas .var(), ddof=1 (tri:494)
        hyper["beta_0"][0, 2] = m                                       # tri:499
        st["beta"] = hyper["beta_0"].copy()
    return st


def sweep(cbs: Cbs, st, hyper, src, D, n_mh_steps, compat="reference"):
    """One Gibbs sweep, in the reference's block order.  Mutates/returns ``st``."""
    N = cbs.N
    st["z"] = draw_z(cbs.t_x, cbs.T_cal, st["lam"], st["mu"], src.u_z(N))
    st["tau"] = draw_tau(cbs.t_x, cbs.T_cal, st["lam"], st["mu"], st["z"],
                         src.e_tau(st["z"]), src.u_tau(~st["z"]))
    if D == 2:                                   # bi:393-399
        Y = np.column_stack([np.log(st["lam"]), np.log(st["mu"])])
        st["beta"], st["Sigma"] = draw_level_2(cbs.X, Y, hyper, src, compat)
        st["lam"], st["mu"] = draw_level_1(cbs, st["lam"], st["mu"], st["z"], st["tau"],
                                           st["beta"], st["Sigma"], src, n_mh_steps)
    else:                                        # tri:519-536
        st["lam"], st["mu"] = draw_level_1(cbs, st["lam"], st["mu"], st["z"], st["tau"],
                                           st["beta"], st["Sigma"], src, n_mh_steps)
        st["eta"] = np.exp(draw_log_eta(cbs.log_s, cbs.X, st["beta"], st["Sigma"],
                                        st["omega2"], src.n_eta(N)))
        Y = np.column_stack([np.log(st["lam"]), np.log(st["mu"]), np.log(st["eta"])])
        st["beta"], st["Sigma"] = draw_level_2(cbs.X, Y, hyper, src, compat)
    return st


def draw_loglik(cbs: Cbs, st):
    """Per-draw mean over customers of the likelihood part (bi:423-428)."""
    lam, mu, z, tau = st["lam"], st["mu"], st["z"], st["tau"]
    lik = cbs.x * np.log(lam) + (1 - z) * np.log(mu) - (lam + mu) * (z * cbs.T_cal + (1 - z) * tau)
    return np.mean(lik)


def level2_row(beta, Sigma):
    """bi:411-412, tri:549-554: beta.T.ravel() then upper triangle of Sigma row-major."""
    D = Sigma.shape[0]
    iu = np.triu_indices(D)
    return np.concatenate([beta.T.ravel(), Sigma[iu]])


def run_chain(cbs: Cbs, hyper, src, mcmc, burnin, thin, D=2, n_mh_steps=20,
              compat="reference"):
    hyper = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in hyper.items()}
    st = init_state(cbs, hyper, D)
    n_draws = (mcmc - 1) // thin + 1
    ncol = 4 if D == 2 else 5
    K = cbs.K
    lvl1 = np.empty((n_draws, cbs.N, ncol))
    lvl2 = np.empty((n_draws, D * K + D * (D + 1) // 2))
    ll = []
    idx = -1
    for step in range(1, burnin + mcmc + 1):
        src.begin_sweep(step)
        sweep(cbs, st, hyper, src, D, n_mh_steps, compat)
        if step > burnin and (step - 1 - burnin) % thin == 0:   # bi:402
            idx += 1
            lvl1[idx, :, 0] = st["lam"]
            lvl1[idx, :, 1] = st["mu"]
            lvl1[idx, :, 2] = st["tau"]
            lvl1[idx, :, 3] = st["z"].astype(float)
            if D == 3:
                lvl1[idx, :, 4] = st["eta"]
            lvl2[idx] = level2_row(st["beta"], st["Sigma"])
            ll.append(draw_loglik(cbs, st))
    return dict(level_1=lvl1, level_2=lvl2, log_likelihood=np.array(ll), state=st)


# ---------------------------------------------------------------------------
# a8  forecast  (bi:506-546, tri:660-749)
# ---------------------------------------------------------------------------
RK_TABLE_SIZE = 64


def poisson_inversion(m, u):
    """Sequential CDF inversion from zero, the injected-stream definition of a
    Poisson draw (the reference calls ``rng.poisson``, NumPy internals; SURVEY
    §8c: "Poisson-from-one-uniform inversion" is a stub definition).
    Arithmetic is fixed so the device matches bit for bit:
    p0 = exp(-m); p_k = (p_{k-1} * m) * rk(k), rk(k) = 1.0/k; stop at first k with
    u <= cdf_k; hard cap at k = 4096 + 16*m."""
    m = np.asarray(m, dtype=float)
    u = np.asarray(u, dtype=float)
    out = np.zeros(m.shape, dtype=np.int64)
    p = np.exp(-m)
    cdf = p.copy()
    active = u > cdf
    k = 0
    cap = np.floor(4096.0 + 16.0 * m)
    while active.any():
        k += 1
        rk = 1.0 / float(k)
        p = np.where(active, (p * m) * rk, p)
        cdf = np.where(active, cdf + p, cdf)
        out = np.where(active, k, out)
        active = active & (u > cdf) & (k < cap)
    return out


def future_horizon(T_cal, tau, z_flag, T_star):
    """bi:535-540: remaining lifetime inside the hold-out window."""
    alive = z_flag > 0.5
    return np.where(alive, T_star, np.clip(tau - T_cal, 0.0, T_star))


def forecast(T_cal, level1_draws, T_star, u, eps=None, sigma_s=0.5):
    """level1_draws: (n_total, N, 4|5); u: (n_total, N) uniforms.
    eps: optional list (per draw) of per-transaction normals in the reference's
    order (customers ascending, each repeated x* times; tri:731-737).
    Returns x* (int64) and, when eps is given, spend (f64)."""
    n_total, N, _ = level1_draws.shape
    xs = np.empty((n_total, N), dtype=np.int64)
    spend = None if eps is None else np.zeros((n_total, N))
    for d in range(n_total):
        lam, tau, zf = level1_draws[d, :, 0], level1_draws[d, :, 2], level1_draws[d, :, 3]
        h = future_horizon(T_cal, tau, zf, T_star)
        xs[d] = poisson_inversion(lam * h, u[d])
        if eps is not None and xs[d].sum() > 0:
            eta = level1_draws[d, :, 4]
            idx = np.repeat(np.arange(N), xs[d])
            per_trx = np.exp(eta[idx] + sigma_s * eps[d][: idx.size])   # tri:732-734 (Q7)
            spend[d] = np.bincount(idx, weights=per_trx, minlength=N)
    return xs if eps is None else (xs, spend)
