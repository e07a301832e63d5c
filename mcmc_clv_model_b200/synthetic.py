"""Synthetic Pareto/NBD customers: `generate_pareto_abe` / `elog2cbs` of bi:75-187, device-generated.

The reference simulates each customer's purchase process in a Python loop (bi:149-162), unusable at the
10 M customers of the headline benchmark.  The device kernel draws the same law without the event loop:
x ~ Poisson(lambda * min(tau, T_cal)), t_x = min(tau, T_cal) * max of x uniforms,
x_star ~ Poisson(lambda * max(0, min(tau, T_cal + T_star) - T_cal)).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib as L

# the C4 benchmark population (CDNOW-like hyper-parameters)
C4_BETA = np.array([[-3.5, -3.6], [0.2, -0.1], [-0.1, 0.1], [0.1, 0.0], [0.0, 0.2]])
C4_GAMMA = np.array([[1.4, 0.3], [0.3, 2.5]])
C4_SEED = 20090601
C4_T_CAL = (27.0, 38.857142857142854)


def generate_cbs_arrays(n, beta, gamma, T_cal=(32.0, 32.0), T_star=39.0, seed=0, gid_offset=0, device=0,
                        X=None, T_cal_values=None, with_truth=True):
    """Raw device generator: dict of NumPy columns x (int32), t_x, T_cal, X (n,K), x_star, lambda, mu, tau."""
    lib = L.load()
    beta = np.ascontiguousarray(beta, dtype=np.float64)
    gamma = np.ascontiguousarray(gamma, dtype=np.float64)
    K = beta.shape[0]
    cfg = L.GenerateConfig(device=device, n_cov=K, n=int(n), gid_offset=int(gid_offset),
                           seed=int(seed) & 0xFFFFFFFFFFFFFFFF, T_cal_lo=float(T_cal[0]), T_cal_hi=float(T_cal[1]),
                           T_star=float(T_star))
    out = dict(x=np.empty(n, dtype=np.int32), t_x=np.empty(n), T_cal=np.empty(n), X=np.empty((n, K)),
               x_star=np.empty(n, dtype=np.int32))
    if X is not None:
        out["X"][:] = np.asarray(X, dtype=np.float64)
    if T_cal_values is not None:
        out["T_cal"][:] = np.asarray(T_cal_values, dtype=np.float64)
    truth = [np.empty(n) for _ in range(3)] if with_truth else [None] * 3
    L.check(lib.clv_generate(C.byref(cfg), L.dptr(beta), L.dptr(gamma), 1 if X is not None else 0,
                             1 if T_cal_values is not None else 0, out["x"].ctypes.data_as(L.c_int32_p),
                             L.dptr(out["t_x"]), L.dptr(out["T_cal"]), L.dptr(out["X"]),
                             out["x_star"].ctypes.data_as(L.c_int32_p), *[L.dptr(t) for t in truth]))
    if with_truth:
        out.update(lambda_true=truth[0], mu_true=truth[1], tau_true=truth[2])
    return out


def elog2cbs(elog, T_cal: float):
    """Event log (cust, t) -> CBS (cust, x, t_x, T_cal); x excludes the first purchase (bi:75-89)."""
    import pandas as pd
    cust = elog["cust"].to_numpy()
    t = elog["t"].to_numpy(float)
    keep = t <= T_cal
    ids, inv = np.unique(cust[keep], return_inverse=True)
    cnt = np.bincount(inv, minlength=ids.size)
    last = np.full(ids.size, -np.inf)
    np.maximum.at(last, inv, t[keep])
    return pd.DataFrame({"cust": ids, "x": np.clip(cnt - 1, 0, None), "t_x": last, "T_cal": T_cal})


def generate_pareto_abe(n: int, T_cal, T_star, beta, gamma, covars: Optional[np.ndarray] = None,
                        seed: Optional[int] = None, *, cbs_clock: str = "reference"):
    """Simulate customers under Abe (2009) -- signature and returned (cbs, elog) frames of bi:95-187.

    The CBS columns come from the device generator; the event log is laid out on the host from the
    generated (x, t_x, x_star, tau): first purchase at t = 0, x - 1 uniform order statistics below t_x,
    t_x itself, and the hold-out purchases uniform on (T_cal, min(tau, T_cal + max T_star)].

    cbs_clock="reference" (default) reproduces the reference's CBS for cohorts with different T_cal: the event log is
    shifted by the birth offsets T_zero = max(T_cal) - T_cal (bi:158) and summarised with ONE calibration end
    (`elog2cbs(elog, T_cal_fix)`, bi:165), so `t_x` is on the shifted clock (T_zero for a customer without repeat
    purchase) and the `T_cal` column is the scalar max(T_cal).  cbs_clock="customer" returns each customer's own clock
    (t_x since the first purchase, per-customer T_cal) -- what the sampler's likelihood actually needs.  The two are
    identical for a scalar T_cal.
    """
    import pandas as pd
    beta = np.asarray(beta, dtype=float)
    K, D = beta.shape
    assert D == 2, "beta must have two columns (log-lambda, log-mu)"          # bi:115
    if seed is None:
        seed = int(np.random.SeedSequence().entropy) & 0x7FFFFFFFFFFFFFFF
    X = None
    if covars is not None:                                                     # bi:123-130
        X = np.asarray(covars, dtype=float)
        if X.ndim == 1:
            X = X[:, None]
        if not np.allclose(X[:, 0], 1):
            X = np.column_stack([np.ones(X.shape[0]), X])
        if X.shape != (n, K):
            raise ValueError("covars has wrong shape relative to beta")
    T_cal = np.asarray(T_cal, dtype=float).ravel()
    if T_cal.size == 1:
        T_cal = np.full(n, T_cal.item())
    T_star = np.asarray(T_star, dtype=float).ravel()
    T_cal_fix = T_cal.max()
    T_zero = T_cal_fix - T_cal
    g = generate_cbs_arrays(n, beta, gamma, T_star=float(T_star.max()), seed=seed, X=X, T_cal_values=T_cal)
    # ---- event log on the host (auxiliary output) -------------------------------------------
    rs = np.random.default_rng(seed)
    x, xs = g["x"].astype(np.int64), g["x_star"].astype(np.int64)
    cust = np.arange(1, n + 1)
    inner = np.maximum(x - 1, 0)
    t_inner = rs.random(int(inner.sum())) * np.repeat(g["t_x"], inner)
    hi = np.minimum(g["tau_true"], T_cal + T_star.max())
    t_hold = np.repeat(T_cal, xs) + rs.random(int(xs.sum())) * np.repeat(hi - T_cal, xs)
    ec = np.concatenate([cust, np.repeat(cust, inner), cust[x > 0], np.repeat(cust, xs)])
    et = np.concatenate([np.zeros(n), t_inner, g["t_x"][x > 0], t_hold])
    et = et + T_zero[ec - 1]
    order = np.lexsort((et, ec))
    elog = pd.DataFrame({"cust": ec[order].astype(float), "t": et[order]})
    if cbs_clock == "reference":
        cbs = pd.DataFrame({"cust": cust.astype(float), "x": x, "t_x": g["t_x"] + T_zero, "T_cal": float(T_cal_fix)})
    elif cbs_clock == "customer":
        cbs = pd.DataFrame({"cust": cust.astype(float), "x": x, "t_x": g["t_x"], "T_cal": T_cal})
    else:
        raise ValueError("cbs_clock must be 'reference' or 'customer'")
    cbs["lambda_true"], cbs["mu_true"], cbs["tau_true"] = g["lambda_true"], g["mu_true"], g["tau_true"]
    cbs["alive_true"] = (T_zero + g["tau_true"]) > T_cal_fix                   # bi:169
    for t_star in T_star:                                                      # bi:172-181
        col = f"x_star{int(t_star)}" if T_star.size > 1 else "x_star"
        if t_star == T_star.max():
            cbs[col] = xs
        else:
            m = (elog["t"].to_numpy() > T_zero[ec[order] - 1] + T_cal[ec[order] - 1]) & \
                (elog["t"].to_numpy() <= T_zero[ec[order] - 1] + T_cal[ec[order] - 1] + t_star)
            cbs[col] = np.bincount(ec[order][m] - 1, minlength=n)
    for j in range(K):                                                         # bi:184-185
        cbs[f"cov{j}"] = g["X"][:, j]
    return cbs, elog
