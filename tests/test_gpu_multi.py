"""Customer-sharded path on >= 2 GPUs (skipped on a 1-GPU box): bit-identical to the single-GPU run."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("D,p2p", [(2, "0"), (3, "0"), (2, "1"), (3, "1")])
def test_two_gpu_sharded_run_is_bit_identical(D, p2p):
    """p2p=1: the all-reduce runs inside k_level2 over peer mailboxes (NVLink P2P stores) instead of NCCL."""
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tools", "sharded_check.py"),
                        "150001", str(D)], capture_output=True, text=True, timeout=600, env=dict(os.environ, CLV_P2P=p2p))
    assert r.returncode == 0 and "SHARDED_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_chains_across_gpus_equal_single_gpu():
    """Chains are independent (bi:485-501): spreading them over two GPUs (one handle and host thread per GPU, no
    communication) gives exactly the chains of the single-GPU call."""
    import numpy as np
    import pandas as pd
    from conftest import load_golden
    from mcmc_clv_model_b200 import mcmc_draw_parameters, mcmc_draw_parameters_rfm_m
    d = load_golden("cdnow_abe.npz")
    cbs = pd.DataFrame({k: d[k][:800] for k in ("x", "t_x", "T_cal", "first_sales_scaled", "log_s")})
    kw = dict(covariates=["first_sales_scaled"], mcmc=12, burnin=20, thin=3, chains=5, seed=7, trace=0)
    for fn in (mcmc_draw_parameters, mcmc_draw_parameters_rfm_m):
        one = fn(cbs, devices=[0], **kw)
        two = fn(cbs, devices=[0, 1], **kw)
        assert len(two["level_1"]) == 5
        for c in range(5):
            np.testing.assert_array_equal(one["level_1"][c], two["level_1"][c])
            np.testing.assert_array_equal(one["level_2"][c], two["level_2"][c])
        assert one["log_likelihood"] == two["log_likelihood"]


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_forecast_sharded_over_devices_equals_single_device():
    """draw_future_transactions with the draws cut into ranges over two GPUs (no collective; the Philox
    counters carry the global draw index) returns exactly the single-GPU x* (and spend)."""
    import numpy as np
    import pandas as pd
    from conftest import load_golden
    from mcmc_clv_model_b200 import draw_future_transactions, draw_future_transactions_rfm_m
    for name, fn in (("fc_bi.npz", draw_future_transactions), ("fc_tri.npz", draw_future_transactions_rfm_m)):
        g = load_golden(name)
        l1 = g["level_1"]
        draws = {"level_1": [l1[:3], l1[3:]]}                      # two "chains" of unequal length
        cbs = pd.DataFrame({"T_cal": g["T_cal"]})
        one = fn(cbs, draws, T_star=39.0, seed=5, devices=[0])
        two = fn(cbs, draws, T_star=39.0, seed=5, devices=[0, 1])
        if isinstance(one, tuple):
            np.testing.assert_array_equal(one[0], two[0])
            np.testing.assert_array_equal(one[1], two[1])
        else:
            np.testing.assert_array_equal(one, two)
