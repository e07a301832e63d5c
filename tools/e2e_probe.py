#!/usr/bin/env python
"""Where does the end-to-end time of one C-ABI call sequence go?  (C4 shape, one GPU)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mcmc_clv_model_b200 import Sampler
from mcmc_clv_model_b200.synthetic import C4_BETA, C4_GAMMA, C4_SEED, C4_T_CAL, generate_cbs_arrays
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 200
c = generate_cbs_arrays(n, C4_BETA, C4_GAMMA, T_cal=C4_T_CAL, seed=C4_SEED, with_truth=False)
pin = {k: torch.from_numpy(np.ascontiguousarray(c[k])).pin_memory().numpy() for k in ("x", "t_x", "T_cal", "X")}
s0 = Sampler(pin["x"], pin["t_x"], pin["T_cal"], pin["X"], chains=1, seed=42)
s0.advance(10); ms = s0.advance_timed(K); s0.close()
print(f"device-resident: {ms:.1f} ms for {K} sweeps")
for it, (kk, pinned) in enumerate([(10, False), (K, False), (K, False), (K, True), (K, True), (K, False)]):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s = Sampler(pin["x"], pin["t_x"], pin["T_cal"], pin["X"], chains=1, seed=42)
    t1 = time.perf_counter()
    out = s.run(0, kk, kk, store_level1=True, pinned=pinned)
    t2 = time.perf_counter()
    s.close()
    t3 = time.perf_counter()
    del out
    t4 = time.perf_counter()
    print(f"it{it} K={kk} pinned={pinned}: create+data+init {t1-t0:.3f}s  run {t2-t1:.3f}s  close {t3-t2:.3f}s  del {t4-t3:.3f}s total {t3-t0:.3f}s -> {n*kk/(t3-t0):.4g}/s", flush=True)
