"""The reference's public entry points, served by the CUDA library.

Same names, arguments, validation, return layout and error behaviour as
`mcmc_draw_parameters` (bi:437-504), `mcmc_draw_parameters_rfm_m` (tri:580-657) and the two
`draw_future_transactions` (bi:506-546, tri:660-749) of lucagem29/mcmc_clv_model, so the reference's
run_mcmc_* / analysis_* scripts work unchanged on the returned (and pickled) dict.

Extra, keyword-only knobs (also settable through the environment so unchanged driver scripts can use
them): rng = "fast" | "strict" (CLV_RNG), compat = "reference" | "paper" (CLV_COMPAT), devices = list of
CUDA ordinals for chains-across-GPUs (CLV_DEVICES="0,1,2,3").
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from typing import Any, Dict, Optional, Sequence

import numpy as np

from . import _lib as L
from .sampler import Sampler, default_hyper

import ctypes as C


def _env_devices():
    s = os.environ.get("CLV_DEVICES")
    if s:
        return [int(t) for t in s.split(",") if t.strip() != ""]
    return [0]


def _seed_value(seed):
    if seed is None:  # OS entropy, as np.random.default_rng(None) (bi:486)
        return int(np.random.SeedSequence().entropy) & 0x7FFFFFFFFFFFFFFF
    return int(seed)


def _design(cal_cbs, covariates, validate):
    """The columns the sampler reads.  The reference copies the frame and prepends an intercept column to build an (N, K)
    matrix (bi:467-470); here the caller's frame is only READ (so it cannot be mutated either) and the covariate columns
    go to the device as they are -- the intercept is implicit in the kernels (clv_set_data_columns)."""
    if covariates is None:
        covariates = []
    covariates = list(covariates)
    if validate:                                                    # bi:461-465 (tri does not validate)
        for col in ("x", "t_x", "T_cal"):
            if col not in cal_cbs:
                raise ValueError(f"cal_cbs missing required column '{col}'")
        if not all(col in cal_cbs for col in covariates):
            raise ValueError("some covariate columns not in cal_cbs")
    return [cal_cbs[c].to_numpy(dtype=float) for c in covariates]   # KeyError for a missing column, as cbs[cols] (tri:617)


def _run(cal_cbs, covariates, mcmc, burnin, thin, chains, seed, trace, n_mh_steps, D, rng, compat, devices,
         hyper=None, return_samplers=False):
    X = _design(cal_cbs, covariates, validate=(D == 2))
    cbs = cal_cbs
    x = cbs["x"].to_numpy()
    t_x = cbs["t_x"].to_numpy(float)
    T_cal = cbs["T_cal"].to_numpy(float)
    log_s = cbs["log_s"].to_numpy(float) if D == 3 else None        # tri:329, 494
    rng = rng or os.environ.get("CLV_RNG", "fast")
    compat = compat or os.environ.get("CLV_COMPAT", "reference")
    devices = list(devices) if devices is not None else _env_devices()
    chains = int(chains)
    seed_v = _seed_value(seed)
    K = len(X) + 1
    hyper = hyper or default_hyper(K, D)
    tot = int(burnin) + int(mcmc)

    # chains-across-GPUs: contiguous groups of chains, one handle per device, no communication
    ndev = max(1, min(len(devices), chains))
    bounds = [chains * i // ndev for i in range(ndev + 1)]
    groups = [(devices[i], bounds[i], bounds[i + 1] - bounds[i]) for i in range(ndev) if bounds[i + 1] > bounds[i]]

    def work(group):
        dev, off, n = group
        s = Sampler(x, t_x, T_cal, X, log_s, model_dim=D, chains=n, chain_offset=off, n_mh_steps=n_mh_steps,
                    seed=seed_v, rng=rng, compat=compat, device=dev, hyper=hyper)

        def progress(step, total):                                   # bi:384-385
            for c in range(n):
                print(f"chain {off + c + 1} | step {step}/{total}")
        out = s.run(burnin, mcmc, thin, store_level1=True, trace=int(trace or 0),
                    progress=progress if trace else None)
        if return_samplers:
            out["sampler"] = s
        else:
            s.close()
        return out

    if len(groups) == 1:
        outs = [work(groups[0])]
    else:
        with ThreadPoolExecutor(len(groups)) as ex:
            outs = list(ex.map(work, groups))
    N = len(x)
    lvl1, lvl2, lls = [], [], []
    for o in outs:
        for c in range(o["level_2"].shape[0]):
            lvl1.append(o["level_1"][c])
            lvl2.append(o["level_2"][c])
            lls.append(o["loglik_sum"][c] / N)                       # per-draw mean over customers, bi:428
    res = dict(level_1=lvl1, level_2=lvl2, log_likelihood=np.mean(np.concatenate(lls)))   # bi:503-504
    if return_samplers:
        res["_samplers"] = [o["sampler"] for o in outs]
    assert tot >= 1
    return res


def mcmc_draw_parameters(cal_cbs, covariates: Sequence[str] | None = None, mcmc: int = 2500, burnin: int = 500,
                         thin: int = 50, chains: int = 2, seed: Optional[int] = None, trace: int = 100,
                         n_mh_steps: int = 20, *, rng=None, compat=None, devices=None, hyper=None) -> Dict[str, Any]:
    """Abe (2009) Gibbs sampler on a calibration CBS -- drop-in for bi:437-504."""
    return _run(cal_cbs, covariates, mcmc, burnin, thin, chains, seed, trace, n_mh_steps, 2, rng, compat, devices, hyper)


def mcmc_draw_parameters_rfm_m(cal_cbs, covariates: Sequence[str] | None = None, mcmc: int = 2500,
                               burnin: int = 500, thin: int = 50, chains: int = 2, seed: Optional[int] = None,
                               trace: int = 100, n_mh_steps: int = 20, *, rng=None, compat=None, devices=None,
                               hyper=None) -> Dict[str, Any]:
    """3-parameter (lambda, mu, eta) RFM-M sampler -- drop-in for tri:580-657."""
    out = _run(cal_cbs, covariates, mcmc, burnin, thin, chains, seed, trace, n_mh_steps, 3, rng, compat, devices, hyper)
    out["log_likelihood"] = float(out["log_likelihood"])                 # tri:652
    return out


# ---------------------------------------------------------------------------------------------
# forecast
# ---------------------------------------------------------------------------------------------
def _forecast(T_cal, level1_list, T_star, seed, simulate_spend, sigma_s, devices=None):
    """x* (and spend) for every (draw, customer) cell, chain-major like bi:530-531.  With several devices the draws of
    every chain are cut into contiguous ranges, one per device ("shard over customers or draws, no
    collective"): the Philox counters carry the GLOBAL draw index (`draw_offset`), so the result does not depend on the
    number of devices."""
    lib = L.load()
    devices = list(devices) if devices else [0]
    T_cal = np.ascontiguousarray(T_cal, dtype=np.float64)
    N = T_cal.size
    ncol = level1_list[0].shape[2]
    seed_v = _seed_value(seed)
    chains = []
    for chain in level1_list:                                            # bi:530-531: chain-major
        a = np.ascontiguousarray(chain, dtype=np.float64)
        if a.shape[1] != N:
            raise ValueError("level_1 draws and cbs disagree on the number of customers")
        chains.append(a)
    n_total = sum(a.shape[0] for a in chains)
    want_spend = bool(simulate_spend and ncol == 5)
    x_future = np.empty((n_total, N), dtype=np.int64)
    spend = np.empty((n_total, N)) if want_spend else None
    # pieces: (device slot, chain array, first draw of the piece within the chain, count, global index of its first draw)
    pieces = [[] for _ in devices]
    off = 0
    for a in chains:
        nd = a.shape[0]
        cuts = [nd * i // len(devices) for i in range(len(devices) + 1)]
        for slot in range(len(devices)):
            if cuts[slot + 1] > cuts[slot]:
                pieces[slot].append((a, cuts[slot], cuts[slot + 1] - cuts[slot], off + cuts[slot]))
        off += nd

    def work(slot):
        for a, d0, cnt, g0 in pieces[slot]:
            cfg = L.ForecastConfig(device=int(devices[slot]), ncol=ncol, n_draws_total=cnt, n_customers=N, gid_offset=0,
                                   draw_offset=g0, T_star=float(T_star), seed=seed_v & 0xFFFFFFFFFFFFFFFF,
                                   simulate_spend=1 if want_spend else 0, reserved=0, sigma_s=float(sigma_s))
            L.check(lib.clv_forecast(C.byref(cfg), L.dptr(a[d0:d0 + cnt]), L.dptr(T_cal),
                                     x_future[g0:g0 + cnt].ctypes.data_as(L.c_int64_p),
                                     L.dptr(spend[g0:g0 + cnt]) if want_spend else None))

    busy = [i for i in range(len(devices)) if pieces[i]]
    if len(busy) <= 1:
        for i in busy:
            work(i)
    else:
        with ThreadPoolExecutor(len(busy)) as ex:
            list(ex.map(work, busy))
    return x_future, spend


def draw_future_transactions(cbs, draws: Dict[str, Any], T_star: float = 39.0, seed: Optional[int] = None, *,
                             devices=None) -> np.ndarray:
    """Simulated x* for each (draw, customer): (n_draws_total, N) int64 -- drop-in for bi:506-546."""
    x, _ = _forecast(cbs["T_cal"].to_numpy(), draws["level_1"], T_star, seed, False, 0.5, devices=devices or _env_devices())
    return x


def draw_future_transactions_rfm_m(cbs, draws: Dict[str, Any], T_star: float = 39.0, *, simulate_spend: bool = True,
                                   sigma_s: float = 0.50, seed: int | None = None, devices=None):
    """Posterior-predictive x* (and log-normal spend) for the RFM-M model -- drop-in for tri:660-749."""
    x, sp = _forecast(cbs["T_cal"].to_numpy(float), draws["level_1"], T_star, seed, simulate_spend, sigma_s,
                      devices=devices or _env_devices())
    if not simulate_spend:
        return x
    return x, sp
