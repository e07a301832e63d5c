#!/usr/bin/env python
"""Kernel time per sweep as the chain evolves (blocks of `blk` sweeps)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcmc_clv_model_b200 import Sampler
from mcmc_clv_model_b200.synthetic import C4_BETA, C4_GAMMA, C4_SEED, C4_T_CAL, generate_cbs_arrays
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
blk = int(sys.argv[2]) if len(sys.argv) > 2 else 100
nblk = int(sys.argv[3]) if len(sys.argv) > 3 else 20
c = generate_cbs_arrays(n, C4_BETA, C4_GAMMA, T_cal=C4_T_CAL, seed=C4_SEED, with_truth=False)
with Sampler(c["x"], c["t_x"], c["T_cal"], c["X"], chains=1, seed=42) as s:
    s.set_timing(True)
    prev = s.get_state(0)
    for b in range(nblk):
        ms = s.advance_timed(blk)
        k, l2, nt = s.kernel_time_ms()
        st = s.get_state(0)
        moved = np.mean(st["log_lambda"] != prev["log_lambda"])
        prev = st
        print(f"sweeps {b*blk+1:5d}-{(b+1)*blk:5d}: k_sweep {k/nt*1e3:8.1f} us  S00={st['Sigma'][0,0]:.3e} S11={st['Sigma'][1,1]:.3e} "
              f"mean ll={st['log_lambda'].mean():.3f} sd ll={st['log_lambda'].std():.3f} sd lm={st['log_mu'].std():.3f} z={st['z'].mean():.3f}", flush=True)
