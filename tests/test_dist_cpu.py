"""Host-side logic of the customer-sharded path over gloo (world_size 2, CPU): shard plan, exact
partition-independent initialisation statistics, and the sharding protocol itself restated on the oracle
(int64 fixed-point level-2 statistics all-reduced each sweep => the sharded chain equals the unsharded one)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mcmc_clv_model_b200.distributed import dist_exact_sum, shard_bounds
from mcmc_clv_model_b200.hostmath import init_statistics


def test_shard_bounds_cover_and_align():
    for n, w in [(10_000_000, 8), (2357, 4), (1000, 3), (5, 8), (23570, 2)]:
        b = shard_bounds(n, w)
        assert b[0][0] == 0 and b[-1][1] == n
        for (lo, hi), (lo2, _) in zip(b[:-1], b[1:]):
            assert hi == lo2 and (lo % 1024 == 0 or lo == n)
        assert all(hi >= lo for lo, hi in b)
        sizes = [hi - lo for lo, hi in b if hi - lo > 0]
        assert max(sizes) - min(sizes) < 2048 or n < 1024 * w


def _data(n=5000, K=3, seed=0):
    g = np.random.default_rng(seed)
    x = g.poisson(1.2, n)
    T = g.uniform(27, 39, n)
    t_x = np.where(x > 0, T * g.random(n), 0.0)
    X = np.column_stack([np.ones(n), g.normal(size=(n, K - 1))])
    return x, t_x, T, X, g.normal(3, 0.7, n)


def _fx_stats(X, Y, c, scale):
    """int64 fixed-point sufficient statistics of one shard (what k_sweep accumulates)."""
    Yc = Y - c
    xty = np.rint((X[:, :, None] * Yc[:, None, :]) * scale).astype(np.int64).sum(axis=0)
    D = Y.shape[1]
    yty = np.array([np.rint(Yc[:, d] * Yc[:, e] * scale).astype(np.int64).sum() for d in range(D) for e in range(d, D)])
    return np.concatenate([xty.ravel(), yty])


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, t_x, T, X, log_s = _data()
    n = x.size
    lo, hi = shard_bounds(n, world)[rank]
    st = init_statistics(x[lo:hi], t_x[lo:hi], T[lo:hi], X[lo:hi], log_s[lo:hi], n, dist_exact_sum())
    # one "sweep" of the protocol: local fixed-point statistics -> all-reduce(sum, int64) -> identical totals everywhere
    g = np.random.default_rng(7)
    Y = g.normal(-3.5, 1.2, (n, 2))
    c = np.array([-3.4, -3.6])
    part = torch.from_numpy(_fx_stats(X[lo:hi], Y[lo:hi], c, 2.0 ** 30))
    dist.all_reduce(part)
    q.put((rank, st, part.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_two_ranks_exact_statistics_and_level2_allreduce():
    world, port = 2, 29533
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    [p.join(timeout=60) for p in ps]
    x, t_x, T, X, log_s = _data()
    whole = init_statistics(x, t_x, T, X, log_s, x.size)
    g = np.random.default_rng(7)
    Y = g.normal(-3.5, 1.2, (x.size, 2))
    tot = _fx_stats(X, Y, np.array([-3.4, -3.6]), 2.0 ** 30)
    for _, st, part in res:
        for k in ("lam_init", "mean_mu_init", "mean_log_s", "omega2", "max_abs_x"):
            assert st[k] == whole[k], k                      # bit-identical for any sharding
        np.testing.assert_array_equal(st["xtx"], whole["xtx"])
        np.testing.assert_array_equal(part, tot)              # integer sums are order/partition independent
    # and the fixed-point statistics reproduce the f64 ones to ~1e-9 relative
    Yc = Y - np.array([-3.4, -3.6])
    np.testing.assert_allclose(tot[:6].reshape(3, 2) / 2.0 ** 30, X.T @ Yc, rtol=1e-7, atol=1e-5)
