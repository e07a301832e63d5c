#!/usr/bin/env python
"""Generate the committed golden fixtures from the REFERENCE ITSELF.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py

Nothing here is imported by the product.  The reference's own functions
(`draw_z`, `draw_tau`, `_draw_level_1`, `_draw_level_2`, `draw_eta`, `_run_chain`,
`mcmc_draw_parameters[_rfm_m]`, `draw_future_transactions`) are executed
unmodified; random numbers are either the reference's real PCG64 streams
("kat_*" fixtures) or injected through a `numpy.random.Generator` subclass that
replays caller-supplied arrays ("inj_*" fixtures).

Outputs (all small, committed):
  cdnow_abe.npz, cdnow_full.npz      CBS columns used by configs C1-C3 (public CDNOW data)
  kat_bi_m1.npz, kat_bi_m2.npz, kat_tri.npz    real-RNG known-answer chains
  inj_bi_k1.npz, inj_bi_k2.npz, inj_tri_k3.npz, inj_edge.npz   injected-stream trajectories
  fc_bi.npz, fc_tri.npz              forecast with injected uniforms / normals
  elog_abe.npz, elog_full.npz        event logs + the reference's elog2cbs output (`make_golden.py cbs` regenerates only these)
(post_*.npz come from make_posterior_golden.py)
"""
import os
import sys

import numpy as np
import pandas as pd

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)

import src.models.bivariate.mcmc as bi          # noqa: E402
import src.models.trivariate.mcmc as tri        # noqa: E402


# ---------------------------------------------------------------------------
# replay generator
# ---------------------------------------------------------------------------
class Replay(np.random.Generator):
    """Replays full-length per-customer variate arrays through the reference's
    block functions.  The harness patches the module-level block functions so
    the stub knows which block is drawing (`phase`) and what z is."""

    def __new__(cls, *a, **k):
        return super().__new__(cls, np.random.PCG64(0))

    def __init__(self, arrays):
        super().__init__(np.random.PCG64(0))
        self.a = arrays
        self.t = -1
        self.phase = None
        self.z = None
        self.log = []

    def next_sweep(self):
        self.t += 1
        self.s_t3 = 0
        self.s_u = 0

    def random(self, size=None, *a, **k):
        if self.phase == "z":
            return self.a["u_z"][self.t].copy()
        if self.phase == "tau":
            return self.a["u_tau"][self.t][~self.z]
        if self.phase == "l1":
            s = self.s_u
            self.s_u += 1
            return self.a["u_acc"][self.t][s].copy()
        raise RuntimeError(f"random() in phase {self.phase}")

    def exponential(self, scale=1.0, size=None):
        assert self.phase == "tau"
        return np.asarray(scale) * self.a["e_tau"][self.t][self.z]

    def standard_t(self, df, size=None):
        assert self.phase == "l1" and df == 3
        s, which = divmod(self.s_t3, 2)
        self.s_t3 += 1
        return self.a["t3_l" if which == 0 else "t3_m"][self.t][s].copy()

    def normal(self, loc=0.0, scale=1.0, size=None):
        if self.phase == "eta":
            return loc + scale * self.a["n_eta"][self.t]
        assert self.phase == "l2"
        return self.a["iw_norm"][self.t].reshape(size)

    def chisquare(self, df, size=None):
        assert self.phase == "l2"
        self.log.append(("chi2_df", np.array(df, dtype=float).ravel().tolist()))
        return self.a["iw_chi2"][self.t].reshape(size)

    def multivariate_normal(self, mean, cov, *a, **k):
        assert self.phase == "l2"
        return np.asarray(mean) + np.linalg.cholesky(cov) @ self.a["beta_norm"][self.t]


def patch_blocks(mod, stub):
    """Wrap the reference module's block functions so the stub sees the phase."""
    orig = {}

    def wrap(name, phase, pre=None, post=None):
        f = getattr(mod, name)
        orig[name] = f

        def g(*args, **kw):
            if pre:
                pre(*args, **kw)
            stub.phase = phase
            out = f(*args, **kw)
            stub.phase = None
            if post:
                post(out)
            return out
        setattr(mod, name, g)

    wrap("draw_z", "z", pre=lambda *a, **k: stub.next_sweep(), post=lambda z: setattr(stub, "z", np.asarray(z)))
    wrap("draw_tau", "tau")
    wrap("_draw_level_1", "l1")
    wrap("_draw_level_2", "l2")
    if hasattr(mod, "draw_eta"):
        wrap("draw_eta", "eta")
    return orig


def unpatch(mod, orig):
    for k, v in orig.items():
        setattr(mod, k, v)


def make_arrays(rng, T, N, S, D, K, nu0):
    nu_n = nu0 + N
    a = dict(
        u_z=rng.random((T, N)), e_tau=rng.standard_exponential((T, N)), u_tau=rng.random((T, N)),
        t3_l=rng.standard_t(3, (T, S, N)), t3_m=rng.standard_t(3, (T, S, N)),
        u_acc=rng.random((T, S, N)),
        iw_norm=rng.standard_normal((T, D * (D - 1) // 2)),
        iw_chi2=np.stack([rng.chisquare(nu_n - D + 1 + i, T) for i in range(D)], axis=1),
        beta_norm=rng.standard_normal((T, D * K)),
    )
    if D == 3:
        a["n_eta"] = rng.standard_normal((T, N))
    return a


def run_injected(mod, cbs_df, covariates, D, T, S, seed):
    """Reference `_run_chain` (unmodified) driven by injected streams for T sweeps, all kept."""
    N = len(cbs_df)
    K = 1 + len(covariates)
    nu0 = (3 + K) if D == 2 else (4 + K)
    arrays = make_arrays(np.random.default_rng(seed), T, N, S, D, K, nu0)
    stub = Replay(arrays)
    orig = patch_blocks(mod, stub)
    try:
        cbs = cbs_df.copy().reset_index(drop=True)
        cbs["intercept"] = 1.0
        cols = ["intercept"] + list(covariates)
        X = cbs[cols].to_numpy(float)
        hyper = dict(beta_0=np.zeros((K, D)), A_0=np.eye(K) * 0.01, nu_00=nu0, gamma_00=nu0 * np.eye(D))
        if D == 2:
            out = mod._run_chain(1, cbs, X, hyper, T, 0, 1, stub, 0, S)
        else:
            out = mod._run_chain(chain_id=1, cbs=cbs, X=X, hyper=hyper, mcmc=T, burnin=0, thin=1,
                                 rng=stub, trace=0, n_mh_steps=S, covariate_cols=cols)
    finally:
        unpatch(mod, orig)
    res = dict(arrays)
    res.update(level_1=out["level_1"], level_2=out["level_2"], loglik=out["log_likelihood"],
               x=cbs["x"].to_numpy().astype(np.int64), t_x=cbs["t_x"].to_numpy(float),
               T_cal=cbs["T_cal"].to_numpy(float), X=X, S=np.int64(S), D=np.int64(D))
    if D == 3:
        res["log_s"] = cbs["log_s"].to_numpy(float)
    chi_dfs = [e[1] for e in stub.log if e[0] == "chi2_df"]
    res["chi2_df_first"] = np.array(chi_dfs[0])
    return res


def add_log_s(df):
    df = df.copy()
    with np.errstate(divide="ignore"):
        df["log_s"] = np.log(df["sales"] / (df["x"] + 1))
    df["log_s"] = df["log_s"].replace(-np.inf, 0.0).fillna(0.0)   # tri/run_mcmc_full.py:60-67
    df["gender_F"] = 1 - df["gender_binary"]                      # tri/run_mcmc_full.py:98-103
    return df


def cbs_npz(df, path):
    cols = ["x", "t_x", "T_cal", "T_star", "x_star", "sales", "first_sales_scaled", "age_scaled",
            "gender_binary", "log_s", "gender_F"]
    np.savez_compressed(path, **{c: df[c].to_numpy() for c in cols})


def main():
    abe = add_log_s(pd.read_csv(f"{REF}/data/processed/cdnow_abeCBS.csv"))
    full = add_log_s(pd.read_csv(f"{REF}/data/processed/cdnow_fullCBS.csv"))
    cbs_npz(abe, f"{HERE}/cdnow_abe.npz")
    cbs_npz(full, f"{HERE}/cdnow_full.npz")

    # ---- real-RNG known-answer chains -------------------
    d = bi.mcmc_draw_parameters(abe, covariates=[], mcmc=100, burnin=100, thin=1, chains=1, seed=42, trace=0)
    print("KAT M1 level_2[-1]:", d["level_2"][0][-1], d["log_likelihood"])
    np.savez_compressed(f"{HERE}/kat_bi_m1.npz", level_2=d["level_2"][0], level_1_last=d["level_1"][0][-1],
                        level_1_c0=d["level_1"][0][:, :8, :], loglik=d["log_likelihood"])
    d = bi.mcmc_draw_parameters(abe, covariates=["first_sales_scaled"], mcmc=40, burnin=20, thin=2,
                                chains=2, seed=7, trace=0)
    np.savez_compressed(f"{HERE}/kat_bi_m2.npz", level_2=np.array(d["level_2"]),
                        level_1_last=np.array([c[-1] for c in d["level_1"]]), loglik=d["log_likelihood"])
    d = tri.mcmc_draw_parameters_rfm_m(abe, covariates=["gender_F", "age_scaled"], mcmc=30, burnin=20, thin=1,
                                       chains=1, seed=11, trace=0)
    np.savez_compressed(f"{HERE}/kat_tri.npz", level_2=d["level_2"][0], level_1_last=d["level_1"][0][-1],
                        loglik=d["log_likelihood"])

    # ---- injected-stream trajectories through the reference's _run_chain --
    sub = abe.iloc[:96]
    np.savez_compressed(f"{HERE}/inj_bi_k1.npz", **run_injected(bi, sub, [], 2, T=6, S=20, seed=101))
    np.savez_compressed(f"{HERE}/inj_bi_k2.npz", **run_injected(bi, sub, ["first_sales_scaled"], 2, T=6, S=20, seed=102))
    np.savez_compressed(f"{HERE}/inj_bi_k4.npz", **run_injected(
        bi, full.iloc[:160], ["first_sales_scaled", "gender_F", "age_scaled"], 2, T=4, S=8, seed=103))
    np.savez_compressed(f"{HERE}/inj_tri_k3.npz", **run_injected(tri, sub, ["gender_F", "age_scaled"], 3, T=6, S=20, seed=104))
    np.savez_compressed(f"{HERE}/inj_tri_k1.npz", **run_injected(tri, sub, [], 3, T=4, S=5, seed=105))

    # adversarial rows: t_x=0, x=0, large x, T_cal-t_x ~ 0, huge rates via covariates
    edge = pd.DataFrame(dict(
        x=[0, 0, 29, 80, 1, 3, 0, 12, 2, 5, 0, 1],
        t_x=[0.0, 0.0, 38.0, 38.85, 38.857, 1e-3, 0.0, 20.0, 27.0, 13.5, 0.0, 0.5],
        T_cal=[27.0, 38.857, 38.857, 38.857, 38.857, 30.0, 31.0, 20.000001, 27.0, 35.0, 38.0, 29.0],
        big=[0.0, 3.0, -3.0, 6.0, -6.0, 9.0, -9.0, 1.0, -1.0, 12.0, -12.0, 0.5],
    ))
    np.savez_compressed(f"{HERE}/inj_edge.npz", **run_injected(bi, edge, ["big"], 2, T=12, S=20, seed=106))

    # ---- forecast with injected uniforms (Poisson = CDF inversion, stub definition) ----
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from oracle.abe_oracle import poisson_inversion

    class FcStub(np.random.Generator):
        def __new__(cls, *a, **k):
            return super().__new__(cls, np.random.PCG64(0))

        def __init__(self, u, eps):
            super().__init__(np.random.PCG64(0))
            self.u, self.eps, self.i = u, eps, -1

        def poisson(self, lam=1.0, size=None):
            self.i += 1
            return poisson_inversion(np.asarray(lam), self.u[self.i])

        def lognormal(self, mean=0.0, sigma=1.0, size=None):
            return np.exp(np.asarray(mean) + sigma * self.eps[self.i][: np.asarray(mean).size])

    g = np.random.default_rng(107)
    inj = np.load(f"{HERE}/inj_bi_k2.npz")
    lvl1 = [inj["level_1"][:3].copy(), inj["level_1"][3:].copy()]
    lvl1[0][0, :4, 0] = [3.0, 9.5, 0.0, 20.0]       # large Poisson means
    u = g.random((6, lvl1[0].shape[1]))
    stub = FcStub(u, None)
    orig_rng = bi.np.random.default_rng
    bi.np.random.default_rng = lambda seed=None: stub
    try:
        xs = bi.draw_future_transactions(pd.DataFrame(dict(T_cal=inj["T_cal"])), dict(level_1=lvl1), T_star=39.0, seed=1)
    finally:
        bi.np.random.default_rng = orig_rng
    np.savez_compressed(f"{HERE}/fc_bi.npz", level_1=np.concatenate(lvl1), T_cal=inj["T_cal"], u=u, x_star=xs, T_star=39.0)

    inj = np.load(f"{HERE}/inj_tri_k3.npz")
    lvl1 = [inj["level_1"].copy()]
    lvl1[0][..., 4] = np.log(lvl1[0][..., 4])       # keep exp(eta) finite: store small "eta" values
    n_tot, N = lvl1[0].shape[:2]
    u = g.random((n_tot, N))
    eps = g.standard_normal((n_tot, 4096))
    stub = FcStub(u, eps)
    tri.np.random.default_rng = lambda seed=None: stub
    try:
        xs, sp = tri.draw_future_transactions(pd.DataFrame(dict(T_cal=inj["T_cal"])), dict(level_1=lvl1),
                                              T_star=39.0, simulate_spend=True, sigma_s=0.5, seed=1)
    finally:
        tri.np.random.default_rng = orig_rng
    np.savez_compressed(f"{HERE}/fc_tri.npz", level_1=lvl1[0], T_cal=inj["T_cal"], u=u, eps=eps, x_star=xs,
                        spend=sp, T_star=39.0, sigma_s=0.5)
    print("golden fixtures written to", HERE)




def make_cbs_golden():
    """Event log -> CBS with the reference's own elog2cbs (src/models/utils/elog2cbs2param.py), as the data-processing
    scripts call it (src/data_processing/2A_cdnow_elog2cbs_abe.py: units="W", T_cal="1997-09-30", T_tot="1998-06-30")."""
    from src.models.utils.elog2cbs2param import elog2cbs
    for name in ("abe", "full"):
        elog = pd.read_csv(f"{REF}/data/raw/cdnow_{name}Elog.csv")
        elog["date"] = pd.to_datetime(elog["date"])
        cbs = elog2cbs(elog, units="W", T_cal="1997-09-30", T_tot="1998-06-30")
        day = ((elog["date"] - pd.Timestamp("1970-01-01")) // pd.Timedelta(days=1)).to_numpy().astype(np.int32)
        np.savez_compressed(f"{HERE}/elog_{name}.npz", cust=elog["cust"].to_numpy().astype(np.int64), day=day,
                            sales=elog["sales"].to_numpy(float),
                            **{f"cbs_{c}": cbs[c].to_numpy() for c in ("cust", "x", "t_x", "litt", "sales", "sales_x", "T_cal",
                                                                      "T_star", "x_star", "sales_star")},
                            cbs_first=((cbs["first"] - pd.Timestamp("1970-01-01")) // pd.Timedelta(days=1)).to_numpy().astype(np.int32))




def make_covariate_golden():
    """Raw customer covariates (data/raw/cdnow_fullCovar.csv) and the standardised columns the reference's data-processing
    script derives from them (src/data_processing/2B_cdnow_elog2cbs_full.py:62-105), as committed in
    data/processed/cdnow_fullCBS.csv:  python tests/golden/make_golden.py covar"""
    cov = pd.read_csv(f"{REF}/data/raw/cdnow_fullCovar.csv")
    cbs = pd.read_csv(f"{REF}/data/processed/cdnow_fullCBS.csv")
    m = cbs[["cust"]].merge(cov, on="cust", how="left")
    assert (m["cust"].to_numpy() == cbs["cust"].to_numpy()).all()
    np.savez_compressed(f"{HERE}/covar_full.npz", cust=cbs["cust"].to_numpy().astype(np.int64),
                        age=m["age"].to_numpy(float), gender_is_M=(m["gender"] == "M").to_numpy().astype(np.int8),
                        gender_is_F=(m["gender"] == "F").to_numpy().astype(np.int8),
                        first_sales_scaled=cbs["first_sales_scaled"].to_numpy(float), age_scaled=cbs["age_scaled"].to_numpy(float),
                        gender_binary=cbs["gender_binary"].to_numpy(float))


if __name__ == "__main__":
    if "covar" in sys.argv[1:]:
        make_covariate_golden()
    else:
        if "cbs" not in sys.argv[1:]:
            main()
        make_cbs_golden()
