"""`draw_z` / `draw_tau` of the reference's `__all__` (bi:193-227, tri:272-304) as device calls.

Both keep the reference's signature `(cbs, lambdas, mus[, z], rng)` and consume the NumPy generator in
the reference's order (N uniforms; n_alive exponentials then n_churn uniforms), but the arithmetic runs
in the CUDA sweep kernel with zero Metropolis steps and injected variates.
"""
from __future__ import annotations

import numpy as np

from .sampler import Sampler


def _one_block(cbs, lambdas, mus, u_z, e_tau, u_tau):
    x = np.zeros(len(cbs), dtype=np.int32)
    t_x = cbs["t_x"].to_numpy(float)
    T_cal = cbs["T_cal"].to_numpy(float)
    N = t_x.size
    X = np.ones((N, 1))
    with Sampler(x, t_x, T_cal, X, model_dim=2, chains=1, n_mh_steps=0, rng="injected",
                 init_stats=dict(lam_init=1.0, mean_mu_init=1.0, mean_log_s=0.0, omega2=1.0, max_abs_x=1.0,
                                 xtx=np.array([[float(N)]]))) as s:
        s.set_state(0, log_lambda=np.log(lambdas), log_mu=np.log(mus))
        s.sweep_injected(dict(u_z=u_z, e_tau=e_tau, u_tau=u_tau, t3_l=np.zeros((1, 0, N)), t3_m=np.zeros((1, 0, N)),
                              u_acc=np.zeros((1, 0, N)), iw_norm=np.zeros((1, 1)), iw_chi2=np.full((1, 2), float(N)),
                              beta_norm=np.zeros((1, 2))), keep=False)
        st = s.get_state(0)
    return st["z"] > 0.5, st["tau"]


def draw_z(cbs, lambdas, mus, rng: np.random.Generator) -> np.ndarray:
    """Alive indicator: rng.random(N) < P(alive | lambda, mu, t_x, T_cal)   (bi:193-200)."""
    lambdas = np.asarray(lambdas, dtype=float)
    mus = np.asarray(mus, dtype=float)
    u = rng.random(lambdas.shape)
    z, _ = _one_block(cbs, lambdas, mus, u, np.ones_like(u), np.full_like(u, 0.5))
    return z


def draw_tau(cbs, lambdas, mus, z, rng: np.random.Generator) -> np.ndarray:
    """Dropout time given z (bi:203-227): alive -> T_cal + Exp(1/mu); churned -> doubly truncated exponential."""
    lambdas = np.asarray(lambdas, dtype=float)
    mus = np.asarray(mus, dtype=float)
    z = np.asarray(z, dtype=bool)
    e = np.ones(z.shape)
    u = np.full(z.shape, 0.5)
    if z.any():
        e[z] = rng.standard_exponential(int(z.sum()))      # rng.exponential(scale) == scale * standard_exponential
    if (~z).any():
        u[~z] = rng.random(int((~z).sum()))
    _, tau = _one_block(cbs, lambdas, mus, np.where(z, -1.0, 2.0), e, u)    # u_z forces the given z
    return tau
