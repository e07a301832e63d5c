#!/usr/bin/env python
"""A/B harness for the sweep kernel: fixed synthetic C4 shard, fixed sweeps, prints kernel ms per sweep.
Usage: CLV_B200_LIB=/path/to/variant.so python tools/kernel_ab.py [n_customers] [sweeps] [chains] [skip]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcmc_clv_model_b200 import Sampler
from mcmc_clv_model_b200.synthetic import C4_BETA, C4_GAMMA, C4_SEED, C4_T_CAL, generate_cbs_arrays

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
chains = int(sys.argv[3]) if len(sys.argv) > 3 else 1
skip = int(sys.argv[4]) if len(sys.argv) > 4 else 10
rng = sys.argv[5] if len(sys.argv) > 5 else "fast"
init = sys.argv[6] if len(sys.argv) > 6 else "reference"     # "truth": start every chain at the generating parameters
c = generate_cbs_arrays(n, C4_BETA, C4_GAMMA, T_cal=C4_T_CAL, seed=C4_SEED, with_truth=True)
mode = os.environ.get("CLV_SWEEP_MODE", "auto")
with Sampler(c["x"], c["t_x"], c["T_cal"], c["X"], chains=chains, seed=42, rng=rng, sweep_mode=mode) as s:
    if init == "truth":
        for ch in range(chains):
            s.set_state(ch, log_lambda=np.log(c["lambda_true"]), log_mu=np.log(c["mu_true"]), beta=C4_BETA, Sigma=C4_GAMMA)
    s.advance(skip)
    per_kernel = os.environ.get("CLV_NO_TIMING") is None
    if per_kernel:
        s.set_timing(True)          # CUDA events around every launch (forces the two-kernel stream path)
    ms = s.advance_timed(sweeps)
    k, l2, nt = s.kernel_time_ms() if per_kernel else (0.0, 0.0, 1)
    st = s.get_state(0)
print(f"lib={os.path.basename(os.environ.get('CLV_B200_LIB','default'))} n={n} chains={chains} rng={rng} init={init} mode={mode} total {ms/sweeps:.4f} ms/sweep  "
      f"k_sweep {k/nt:.4f} ms  k_level2 {l2/nt*1e3:.1f} us  -> {n*chains*sweeps/(ms*1e-3):.4g} cust-upd/s  S00={st['Sigma'][0,0]:.4g}")
