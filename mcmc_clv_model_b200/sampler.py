"""Host-side mirror of the reference's chain driver over the C-ABI (include/clv_b200.h).

`Sampler` owns one `clv_sampler` handle: one shard of customers x a group of chains on one GPU.
It does what `_run_chain` (bi:346-431, tri:465-574) does around the Gibbs blocks -- validation,
hyper-parameters, initial state, burn-in/thinning bookkeeping, draw storage -- and nothing numerical:
every draw happens in the CUDA library.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib as L
from .hostmath import ExactSum, init_statistics

_RNG = {"fast": L.RNG_FAST, "strict": L.RNG_STRICT, "injected": L.RNG_INJECTED}
_COMPAT = {"reference": L.COMPAT_REFERENCE, "paper": L.COMPAT_PAPER}
_SWEEP = {"auto": L.SWEEP_AUTO, "stream": L.SWEEP_STREAM, "persistent": L.SWEEP_PERSISTENT}


def default_hyper(K: int, D: int):
    """Diffuse NIW prior of bi:474-479 (D=2) / tri:622-626 (D=3)."""
    nu0 = (3 + K) if D == 2 else (4 + K)
    return dict(beta_0=np.zeros((K, D)), A_0=np.eye(K) * 0.01, nu_00=float(nu0), gamma_00=nu0 * np.eye(D))


def _pinned_empty(shape):
    """Page-locked float64 host array (torch is used for the allocation only)."""
    try:
        import torch
        if torch.cuda.is_available():
            return torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()
    except Exception:  # pragma: no cover - allocation plumbing only
        pass
    return np.empty(shape, dtype=np.float64)


class Sampler:
    def __init__(self, x, t_x, T_cal, X, log_s=None, *, model_dim=2, chains=1, chain_offset=0, n_mh_steps=20,
                 seed=0, rng="fast", compat="reference", device=0, n_global=None, gid_offset=0, hyper=None,
                 esum: Optional[ExactSum] = None, sweep_mode="auto", init_stats=None, comm=None):
        """init_stats: None -> exact statistics computed on the device (all-reduced over `comm` when sharded);
        "host" -> hostmath.init_statistics (with `esum` for sharded runs); or an explicit dict.
        comm: (nccl_unique_id_bytes, rank, world) for customer-sharded runs."""
        self.lib = L.load()
        self.h = C.c_void_p()
        x = np.ascontiguousarray(x, dtype=np.int32)
        t_x = np.ascontiguousarray(t_x, dtype=np.float64)
        T_cal = np.ascontiguousarray(T_cal, dtype=np.float64)
        cov_cols = None
        if isinstance(X, (list, tuple)):
            # the covariate columns themselves, as a DataFrame holds them (the intercept of bi:468-470 is implicit): they go
            # to the device one by one (clv_set_data_columns); no (N, K) matrix is assembled on the host
            cov_cols = [np.ascontiguousarray(c, dtype=np.float64) for c in X]
            if any(c.ndim != 1 or c.size != x.size for c in cov_cols):
                raise ValueError("every covariate column must have N entries")
            N, K = x.size, len(cov_cols) + 1
            X = None
        else:
            X = np.ascontiguousarray(X, dtype=np.float64)
            if X.ndim != 2 or X.shape[0] != x.size:
                raise ValueError("X must be (N, K)")
            N, K = X.shape
        D = int(model_dim)
        if D == 3:
            if log_s is None:
                raise ValueError("the trivariate model needs log_s")
            log_s = np.ascontiguousarray(log_s, dtype=np.float64)
        else:
            log_s = None
        self.N, self.K, self.D, self.chains = N, K, D, int(chains)
        self.ncol = 4 if D == 2 else 5
        self.P = D * K + D * (D + 1) // 2
        self.n_global = int(n_global if n_global is not None else N)
        self.S = int(n_mh_steps)
        cfg = L.Config(model_dim=D, n_cov=K, n_chains=self.chains, chain_offset=int(chain_offset),
                       n_mh_steps=self.S, rng_mode=_RNG[rng], compat=_COMPAT[compat], sweep_mode=_SWEEP[sweep_mode],
                       device=int(device), reserved=0, n_local=N, n_global=self.n_global,
                       gid_offset=int(gid_offset), seed=int(seed) & 0xFFFFFFFFFFFFFFFF)
        L.check(self.lib.clv_create(C.byref(self.h), C.byref(cfg)))
        try:
            if cov_cols is not None:
                ptrs = (L.c_double_p * max(len(cov_cols), 1))(*[L.dptr(c) for c in cov_cols])
                L.check(self.lib.clv_set_data_columns(self.h, x.ctypes.data_as(L.c_int32_p), L.dptr(t_x), L.dptr(T_cal),
                                                      ptrs, L.dptr(log_s)), self.h)
                if init_stats == "host" or esum is not None:            # the host statistics take the matrix
                    X = np.column_stack([np.ones(N)] + cov_cols)
            else:
                try:
                    L.check(self.lib.clv_set_data(self.h, x.ctypes.data_as(L.c_int32_p), L.dptr(t_x), L.dptr(T_cal),
                                                  L.dptr(X), L.dptr(log_s)), self.h)
                except L.ClvError as e:        # the intercept column is validated on the device during the upload
                    if "intercept" in str(e):
                        raise ValueError("column 0 of X must be the intercept (all ones)") from None
                    raise
            hy = hyper or default_hyper(K, D)
            b0 = np.ascontiguousarray(hy["beta_0"], dtype=np.float64)
            a0 = np.ascontiguousarray(hy["A_0"], dtype=np.float64)
            g0 = np.ascontiguousarray(hy["gamma_00"], dtype=np.float64)
            if b0.shape != (K, D) or a0.shape != (K, K) or g0.shape != (D, D):
                raise ValueError("hyper-parameter shapes do not match (K, D)")
            L.check(self.lib.clv_set_hyper(self.h, L.dptr(b0), L.dptr(a0), float(hy["nu_00"]), L.dptr(g0)), self.h)
            if comm is not None:
                self.comm_init(*comm)
            if init_stats is None and esum is None:
                L.check(self.lib.clv_init_state(self.h, None), self.h)
                cst, xtx = L.InitStats(), np.empty((K, K))
                L.check(self.lib.clv_get_init_stats(self.h, C.byref(cst), L.dptr(xtx)), self.h)
                self.init_stats = dict(lam_init=cst.lam_init, mean_mu_init=cst.mean_mu_init, mean_log_s=cst.mean_log_s,
                                       omega2=cst.omega2, max_abs_x=cst.max_abs_x, xtx=xtx)
            else:
                st = init_stats if isinstance(init_stats, dict) else init_statistics(x, t_x, T_cal, X, log_s, self.n_global, esum)
                self.init_stats = st
                xtx = np.ascontiguousarray(st["xtx"], dtype=np.float64)
                cst = L.InitStats(lam_init=st["lam_init"], mean_mu_init=st["mean_mu_init"], mean_log_s=st["mean_log_s"],
                                  omega2=st["omega2"], max_abs_x=st["max_abs_x"], xtx=L.dptr(xtx))
                L.check(self.lib.clv_init_state(self.h, C.byref(cst)), self.h)
        except Exception:
            self.close()
            raise

    # ---- lifecycle ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.clv_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- communication -----------------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        L.check(L.load().clv_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        buf = C.create_string_buffer(unique_id, 128)
        L.check(self.lib.clv_comm_init(self.h, buf, int(rank), int(world)), self.h)

    def p2p_export(self) -> bytes:
        buf = C.create_string_buffer(64)
        L.check(self.lib.clv_p2p_export(self.h, buf), self.h)
        return buf.raw

    def p2p_is_cached(self, rank: int, world: int) -> bool:
        return bool(self.lib.clv_p2p_is_cached(self.h, int(rank), int(world)))

    def p2p_connect(self, handles, rank: int, world: int):
        """handles: list of the `world` 64-byte mailbox handles (p2p_export of every rank, all-gathered); None when
        p2p_is_cached()."""
        blob = C.create_string_buffer(b"".join(handles), 64 * world) if handles is not None else None
        L.check(self.lib.clv_p2p_connect(self.h, blob, int(rank), int(world)), self.h)

    # ---- sweeps ------------------------------------------------------------------------------
    def run(self, burnin, mcmc, thin, store_level1=True, trace=0, progress=None, pinned=False, out=None):
        """burnin + mcmc sweeps; returns dict(level_1 [chains] of (n_draws,N,ncol) | None,
        level_2 (chains,n_draws,P), loglik_sum (chains,n_draws) = per-draw SUM over local customers).
        pinned=True page-locks the level-1 output (worth it only when the draws do not fit one device chunk and
        are streamed out while the sweeps continue).  out: a caller-provided C-contiguous float64 array of shape
        (chains, n_draws, N, ncol) for the level-1 draws (e.g. one reused over several runs)."""
        n_draws = (int(mcmc) - 1) // int(thin) + 1
        shape = (self.chains, n_draws, self.N, self.ncol)
        if out is not None:
            if not store_level1 or out.shape != shape or out.dtype != np.float64 or not out.flags.c_contiguous:
                raise ValueError(f"out must be a C-contiguous float64 array of shape {shape}")
            lvl1 = out
        else:
            lvl1 = (_pinned_empty(shape) if pinned else np.empty(shape)) if store_level1 else None
        lvl2 = np.empty((self.chains, n_draws, self.P))
        ll = np.empty((self.chains, n_draws))
        cb = L.PROGRESS_CB(lambda user, step, total: progress(int(step), int(total))) if progress else C.cast(None, L.PROGRESS_CB)
        L.check(self.lib.clv_run(self.h, int(burnin), int(mcmc), int(thin), L.dptr(lvl1) if lvl1 is not None else None,
                                 L.dptr(lvl2), L.dptr(ll), cb, None, int(trace)), self.h)
        return dict(level_1=lvl1, level_2=lvl2, loglik_sum=ll)

    def run_resident(self, burnin, mcmc, thin, trace=0, progress=None):
        """Like run(), but the level-1 draws stay in HBM (for forecast_resident / zero-copy consumers)."""
        n_draws = (int(mcmc) - 1) // int(thin) + 1
        lvl2 = np.empty((self.chains, n_draws, self.P))
        ll = np.empty((self.chains, n_draws))
        cb = L.PROGRESS_CB(lambda user, step, total: progress(int(step), int(total))) if progress else C.cast(None, L.PROGRESS_CB)
        L.check(self.lib.clv_run_resident(self.h, int(burnin), int(mcmc), int(thin), L.dptr(lvl2), L.dptr(ll), cb, None,
                                          int(trace)), self.h)
        self._resident = n_draws
        return dict(level_1=None, level_2=lvl2, loglik_sum=ll)

    def advance(self, n_sweeps, sync=True):
        L.check(self.lib.clv_advance(self.h, int(n_sweeps), 1 if sync else 0), self.h)

    def advance_timed(self, n_sweeps) -> float:
        """n sweeps, synchronous; returns their device time in ms (CUDA events on the handle's stream)."""
        ms = C.c_double()
        L.check(self.lib.clv_advance_timed(self.h, int(n_sweeps), C.byref(ms)), self.h)
        return ms.value

    @staticmethod
    def lockstep_advance(shards, n_sweeps):
        """Test hook (clv_debug_lockstep_advance): advance customer shards that live on ONE device in lockstep, the
        level-2 all-reduce running the production peer-mailbox protocol with the ranks emulated inside one kernel."""
        arr = (C.c_void_p * len(shards))(*[s.h for s in shards])
        L.check(shards[0].lib.clv_debug_lockstep_advance(arr, len(shards), int(n_sweeps)), shards[0].h)

    @property
    def sweeps_done(self):
        return int(self.lib.clv_sweeps_done(self.h))

    @sweeps_done.setter
    def sweeps_done(self, n):
        L.check(self.lib.clv_set_sweeps_done(self.h, int(n)), self.h)

    def checkpoint(self):
        """Everything needed to continue the chains in another process: per-chain state + the sweep counter."""
        return dict(sweeps_done=self.sweeps_done, chains=[self.get_state(c) for c in range(self.chains)])

    def restore(self, ck):
        """Counterpart of checkpoint() on a sampler built from the same data, seed and options."""
        for c, st in enumerate(ck["chains"]):
            self.set_state(c, st["log_lambda"], st["log_mu"], st.get("log_eta"), st["beta"], st["Sigma"])
        self.sweeps_done = ck["sweeps_done"]

    @property
    def kernel_launches(self):
        return int(self.lib.clv_kernel_launches(self.h))

    def set_timing(self, on=True):
        L.check(self.lib.clv_set_timing(self.h, 1 if on else 0), self.h)

    def kernel_time_ms(self):
        a, b, n = C.c_double(), C.c_double(), C.c_int64()
        L.check(self.lib.clv_kernel_time_ms(self.h, C.byref(a), C.byref(b), C.byref(n)), self.h)
        return a.value, b.value, n.value

    # ---- state -------------------------------------------------------------------------------
    def get_state(self, chain=0):
        N, K, D = self.N, self.K, self.D
        out = dict(log_lambda=np.empty(N), log_mu=np.empty(N), z=np.empty(N), tau=np.empty(N),
                   beta=np.empty((K, D)), Sigma=np.empty((D, D)))
        le = np.empty(N) if D == 3 else None
        L.check(self.lib.clv_get_state(self.h, int(chain), L.dptr(out["log_lambda"]), L.dptr(out["log_mu"]), L.dptr(le),
                                       L.dptr(out["z"]), L.dptr(out["tau"]), L.dptr(out["beta"]), L.dptr(out["Sigma"])),
                self.h)
        if D == 3:
            out["log_eta"] = le
        return out

    def set_state(self, chain=0, log_lambda=None, log_mu=None, log_eta=None, beta=None, Sigma=None):
        f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
        arrs = [f(a) for a in (log_lambda, log_mu, log_eta, beta, Sigma)]
        L.check(self.lib.clv_set_state(self.h, int(chain), *[L.dptr(a) for a in arrs]), self.h)

    def sweep_injected(self, v: dict, keep=True):
        """One sweep of every chain with caller-supplied variates (arrays chain-major, see clv_injected)."""
        C_, N, S, D, K = self.chains, self.N, self.S, self.D, self.K
        shapes = dict(u_z=(C_, N), e_tau=(C_, N), u_tau=(C_, N), t3_l=(C_, S, N), t3_m=(C_, S, N), u_acc=(C_, S, N),
                      n_eta=(C_, N), iw_norm=(C_, D * (D - 1) // 2), iw_chi2=(C_, D), beta_norm=(C_, D * K))
        keepalive, inj = [], L.Injected()
        for name, shp in shapes.items():
            if name == "n_eta" and D == 2:
                continue
            a = np.ascontiguousarray(np.asarray(v[name], dtype=np.float64).reshape(shp))
            keepalive.append(a)
            setattr(inj, name, L.dptr(a))
        lvl1 = np.empty((C_, N, self.ncol)) if keep else None
        lvl2 = np.empty((C_, self.P)) if keep else None
        ll = np.empty(C_) if keep else None
        L.check(self.lib.clv_sweep_injected(self.h, C.byref(inj), 1 if keep else 0, L.dptr(lvl1), L.dptr(lvl2), L.dptr(ll)),
                self.h)
        return dict(level_1=lvl1, level_2=lvl2, loglik_sum=ll)

    # ---- fused forecast --------------------------------------------------------------------------
    def set_fused_forecast(self, T_star=39.0, seed=0, enable=True):
        """Simulate x* of every kept draw inside the sweep kernel (lambda, tau, z still in registers) during the next
        run()/run_resident(): no pass over stored draws, which need not even be kept (store_level1=False)."""
        L.check(self.lib.clv_set_fused_forecast(self.h, 1 if enable else 0, float(T_star), int(seed) & 0xFFFFFFFFFFFFFFFF), self.h)

    def fused_forecast_result(self):
        """Per-customer mean x* and P(alive) over all kept draws of all chains of the last run (same values as
        forecast_resident() on the same draws and seed)."""
        mx, pa, n = np.empty(self.N), np.empty(self.N), C.c_int64()
        L.check(self.lib.clv_fused_forecast_result(self.h, L.dptr(mx), L.dptr(pa), C.byref(n)), self.h)
        return dict(mean_x_star=mx, p_alive=pa, n_draws_total=int(n.value))

    # ---- resident forecast ---------------------------------------------------------------------
    def forecast_resident(self, T_star=39.0, seed=0, want_x_star=False):
        """x* / P(alive) from the draws still in HBM after run()/run_resident(): per-customer mean x* and mean z,
        optionally the full (chains*n_draws, N) x* matrix."""
        ptr, nd = C.c_void_p(), C.c_int64()
        L.check(self.lib.clv_resident_draws(self.h, C.byref(ptr), C.byref(nd)), self.h)
        mx, pa, ms = np.empty(self.N), np.empty(self.N), C.c_double()
        xs = np.empty((self.chains * nd.value, self.N), dtype=np.int64) if want_x_star else None
        L.check(self.lib.clv_forecast_resident(self.h, float(T_star), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                               xs.ctypes.data_as(L.c_int64_p) if xs is not None else None,
                                               L.dptr(mx), L.dptr(pa), C.byref(ms)), self.h)
        return dict(mean_x_star=mx, p_alive=pa, x_star=xs, kernel_ms=ms.value, n_draws_total=self.chains * nd.value)

    # ---- analysis reductions on the resident draws -------------------------------------
    SUMMARY_COLUMNS = ("mean_lambda", "lambda_2.5", "lambda_97.5", "mean_mu_capped", "mu_2.5", "mu_97.5", "p_alive",
                       "mean_tau", "mean_mu", "mean_eta")

    def posterior_summary(self, mu_cap=0.05):
        """Per-customer posterior means / percentiles over every resident draw (compute_table4's inputs,
        utils/analysis_bi_helpers.py:75-110) -> dict of (N,) arrays keyed by SUMMARY_COLUMNS."""
        out = np.empty((self.N, len(self.SUMMARY_COLUMNS)))
        L.check(self.lib.clv_posterior_summary(self.h, float(mu_cap), L.dptr(out)), self.h)
        return {k: out[:, j].copy() for j, k in enumerate(self.SUMMARY_COLUMNS)}

    def weekly_tracking(self, birth_week, times, seed=0):
        """Mean over resident draws of the weekly incremental repeat transactions (Figure 2,
        bivariate/analysis_abe.py:446-464); np.cumsum of the result is the tracking curve."""
        b = np.ascontiguousarray(birth_week, dtype=np.float64)
        t = np.ascontiguousarray(times, dtype=np.float64)
        out = np.empty(t.size)
        L.check(self.lib.clv_weekly_tracking(self.h, L.dptr(b), L.dptr(t), int(t.size), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                             L.dptr(out)), self.h)
        return out
