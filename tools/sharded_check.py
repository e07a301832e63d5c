#!/usr/bin/env python
"""torchrun --nproc-per-node G tools/sharded_check.py : customer-sharded run over G GPUs (NCCL all-reduce of the
int64 level-2 statistics every sweep) must reproduce the single-GPU run of the same problem BIT FOR BIT."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from mcmc_clv_model_b200 import Sampler
from mcmc_clv_model_b200.distributed import broadcast_unique_id, gather_level1, shard_bounds
from mcmc_clv_model_b200.synthetic import C4_BETA, C4_GAMMA, C4_SEED, C4_T_CAL, generate_cbs_arrays

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200_003
D = int(sys.argv[2]) if len(sys.argv) > 2 else 2
lo, hi = shard_bounds(N, world)[rank]
c = generate_cbs_arrays(hi - lo, C4_BETA, C4_GAMMA, T_cal=C4_T_CAL, seed=C4_SEED, gid_offset=lo, device=local, with_truth=False)
log_s = (0.5 * c["X"][:, 1] + 3.0 + 0.1 * np.cos(np.arange(lo, hi))) if D == 3 else None
comm = (broadcast_unique_id(Sampler.comm_unique_id), rank, world)
P2P = os.environ.get("CLV_P2P") == "1"
with Sampler(c["x"], c["t_x"], c["T_cal"], c["X"], log_s, model_dim=D, chains=2, seed=9, device=local, n_global=N, gid_offset=lo,
             comm=comm) as s:
    if P2P:
        from mcmc_clv_model_b200.distributed import connect_p2p
        connect_p2p(s)
    out = s.run(5, 6, 2)
    stats = s.init_stats
l1 = gather_level1(np.ascontiguousarray(np.moveaxis(out["level_1"], 0, 1).reshape(out["level_1"].shape[1], -1, 1)))  # noqa
ll = torch.tensor(out["loglik_sum"], device="cuda")
dist.all_reduce(ll)
if rank == 0:
    full = generate_cbs_arrays(N, C4_BETA, C4_GAMMA, T_cal=C4_T_CAL, seed=C4_SEED, gid_offset=0, device=local, with_truth=False)
    ls = (0.5 * full["X"][:, 1] + 3.0 + 0.1 * np.cos(np.arange(N))) if D == 3 else None
    with Sampler(full["x"], full["t_x"], full["T_cal"], full["X"], ls, model_dim=D, chains=2, seed=9, device=local) as s1:
        ref = s1.run(5, 6, 2)
        st1 = s1.init_stats
    for k in ("lam_init", "mean_mu_init", "mean_log_s", "omega2", "max_abs_x"):
        assert stats[k] == st1[k], (k, stats[k], st1[k])
    assert np.array_equal(stats["xtx"], st1["xtx"])
    assert np.array_equal(out["level_2"], ref["level_2"]), np.abs(out["level_2"] - ref["level_2"]).max()
    assert np.array_equal(out["level_1"], ref["level_1"][:, :, lo:hi, :])
    assert np.array_equal(ll.cpu().numpy(), ref["loglik_sum"])
    print(f"SHARDED_OK world={world} N={N} D={D} p2p={P2P}: level_2, level_1, loglik and init statistics bit-identical to the 1-GPU run")
dist.barrier()
dist.destroy_process_group()
