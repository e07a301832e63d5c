#!/usr/bin/env python
"""Stage timing of the C-ABI call sequence behind Sampler(...) for the C4 shape (one GPU): where the set-up time goes."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mcmc_clv_model_b200 import _lib as L
from mcmc_clv_model_b200.sampler import default_hyper
from mcmc_clv_model_b200.synthetic import C4_BETA, C4_GAMMA, C4_SEED, C4_T_CAL, generate_cbs_arrays
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
c = generate_cbs_arrays(n, C4_BETA, C4_GAMMA, T_cal=C4_T_CAL, seed=C4_SEED, with_truth=False)
pin = {k: torch.from_numpy(np.ascontiguousarray(c[k])).pin_memory().numpy() for k in ("x", "t_x", "T_cal", "X")}
lib = L.load()
K = pin["X"].shape[1]
# the covariate columns on their own (what mcmc_draw_parameters hands over: clv_set_data_columns)
c["cov"] = [np.ascontiguousarray(c["X"][:, k]) for k in range(1, K)]
pin["cov"] = [torch.from_numpy(v).pin_memory().numpy() for v in c["cov"]]
for it in range(8):
    src = pin if it % 2 == 0 else c
    columns = it >= 4
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    h = C.c_void_p()
    cfg = L.Config(model_dim=2, n_cov=K, n_chains=1, chain_offset=0, n_mh_steps=20, rng_mode=0, compat=0, sweep_mode=0,
                   device=0, reserved=0, n_local=n, n_global=n, gid_offset=0, seed=42)
    L.check(lib.clv_create(C.byref(h), C.byref(cfg))); t.append(time.perf_counter())
    if columns:
        ptrs = (L.c_double_p * (K - 1))(*[L.dptr(v) for v in src["cov"]])
        L.check(lib.clv_set_data_columns(h, src["x"].ctypes.data_as(L.c_int32_p), L.dptr(src["t_x"]), L.dptr(src["T_cal"]), ptrs, None), h)
    else:
        L.check(lib.clv_set_data(h, src["x"].ctypes.data_as(L.c_int32_p), L.dptr(src["t_x"]), L.dptr(src["T_cal"]), L.dptr(src["X"]), None), h)
    t.append(time.perf_counter())
    hy = default_hyper(K, 2)
    L.check(lib.clv_set_hyper(h, L.dptr(hy["beta_0"]), L.dptr(hy["A_0"]), float(hy["nu_00"]), L.dptr(hy["gamma_00"])), h)
    t.append(time.perf_counter())
    L.check(lib.clv_init_state(h, None), h); t.append(time.perf_counter())
    lib.clv_destroy(h); t.append(time.perf_counter())
    names = ["clv_create", "clv_set_data_columns" if columns else "clv_set_data (N x K matrix)", "clv_set_hyper", "clv_init_state", "clv_destroy"]
    print(("pinned  " if it % 2 == 0 else "pageable") + "  " + "  ".join(f"{nm} {1e3*(b-a):.1f} ms" for nm, a, b in zip(names, t, t[1:])), flush=True)
