"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol include/clv_b200.h
declares, the ctypes mirror covers them, and the product fails loudly without a CUDA device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from mcmc_clv_model_b200 import _lib as L

HEADER = os.path.join(ROOT, "include", "clv_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(clv_[a-z0-9_]+)\s*\(", src))
    names -= {"clv_progress_cb"}
    return names


def test_header_symbols_exported():
    out = subprocess.run(["nm", "-D", "--defined-only", L.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    missing = _declared() - exported
    assert not missing, f"declared in clv_b200.h but not exported: {sorted(missing)}"


def test_ctypes_mirror_is_complete():
    assert _declared() == set(L.SIGNATURES), (_declared() ^ set(L.SIGNATURES))
    lib = L.load()
    assert lib.clv_abi_version() == 2


def test_struct_layouts():
    assert C.sizeof(L.Config) == 10 * 4 + 4 * 8
    assert C.sizeof(L.InitStats) == 6 * 8
    assert C.sizeof(L.Injected) == 10 * 8
    assert C.sizeof(L.ForecastConfig) == 2 * 4 + 4 * 8 + 8 + 8 + 2 * 4 + 8
    assert C.sizeof(L.GenerateConfig) == 2 * 4 + 3 * 8 + 3 * 8


def test_argument_validation_without_device():
    lib = L.load()
    h = C.c_void_p()
    bad = L.Config(model_dim=4, n_cov=1, n_chains=1, n_mh_steps=20, n_local=4, n_global=4)
    assert lib.clv_create(C.byref(h), C.byref(bad)) == -1
    assert b"model_dim" in lib.clv_last_error(None)
    bad = L.Config(model_dim=2, n_cov=99, n_chains=1, n_mh_steps=20, n_local=4, n_global=4)
    assert lib.clv_create(C.byref(h), C.byref(bad)) == -1


def test_column_wise_design_is_validated_on_the_host():
    """Covariate columns (clv_set_data_columns): wrong lengths are refused before any device work, and a null handle /
    null column comes back as an error code, not a crash."""
    import numpy as np
    from mcmc_clv_model_b200 import Sampler
    x, t, T = np.array([0, 1, 2]), np.array([0.0, 3.0, 9.0]), np.full(3, 30.0)
    with pytest.raises(ValueError, match="covariate column"):
        Sampler(x, t, T, [np.zeros(2)])
    with pytest.raises(ValueError, match="covariate column"):
        Sampler(x, t, T, [np.zeros((3, 1))])
    lib = L.load()
    assert lib.clv_set_data_columns(None, None, None, None, None, None) == -1


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device failure mode")
def test_no_cpu_fallback():
    """Without a CUDA device the product raises instead of computing anything on the host."""
    import pandas as pd
    from mcmc_clv_model_b200 import mcmc_draw_parameters
    cbs = pd.DataFrame(dict(x=[0, 1, 2], t_x=[0.0, 3.0, 9.0], T_cal=[30.0, 30.0, 30.0]))
    with pytest.raises(L.ClvError) as e:
        mcmc_draw_parameters(cbs, mcmc=2, burnin=1, thin=1, chains=1, seed=1, trace=0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_reference_validation_errors():
    """Same ValueErrors as bi:461-465, raised before any device work."""
    import pandas as pd
    from src.models.bivariate.mcmc import mcmc_draw_parameters
    with pytest.raises(ValueError, match="missing required column 't_x'"):
        mcmc_draw_parameters(pd.DataFrame(dict(x=[1], T_cal=[3.0])), mcmc=1, burnin=0, thin=1, chains=1)
    with pytest.raises(ValueError, match="some covariate columns not in cal_cbs"):
        mcmc_draw_parameters(pd.DataFrame(dict(x=[1], t_x=[1.0], T_cal=[3.0])), covariates=["nope"])


def test_product_does_not_import_oracle():
    """The shipped package must not reference oracle/ (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "mcmc_clv_model_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
    for f in ("src/models/bivariate/mcmc.py", "src/models/trivariate/mcmc.py"):
        assert "oracle" not in open(os.path.join(ROOT, f)).read().replace("oracle/philox_np.py", "")


def test_exact_sum_is_partition_independent():
    from mcmc_clv_model_b200.hostmath import ExactSum, exact_partial, fx_bits, from_fx, init_statistics
    rng = np.random.default_rng(3)
    v = rng.lognormal(2.0, 2.0, 100_003)
    whole = ExactSum()(v)
    m = float(np.abs(v).max())
    for cuts in ([0, 17, 50_000, 100_003], [0, 99_999, 100_003], list(range(0, 100_004, 10_000)) + [100_003]):
        parts = [v[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
        bits = fx_bits(m)
        tot = sum(exact_partial(p, bits) for p in parts)
        assert from_fx(tot, bits) == whole
    assert abs(whole - np.sum(v)) <= 1e-12 * np.sum(v)
    x = rng.poisson(1.0, 5000)
    t = rng.random(5000) * 30 * (x > 0)
    T = np.full(5000, 38.0)
    X = np.column_stack([np.ones(5000), rng.normal(size=5000)])
    st = init_statistics(x, t, T, X, None, 5000)
    lam = x.mean() / np.mean(np.where(t == 0, T, t))
    assert abs(st["lam_init"] / lam - 1) < 1e-13
    assert abs(st["mean_mu_init"] / np.mean(1 / (t + 0.5 / lam)) - 1) < 1e-13
    np.testing.assert_allclose(st["xtx"], X.T @ X, rtol=1e-13)


def test_missing_library_fails_loudly(monkeypatch):
    """No silent fallback when the CUDA library has not been built."""
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", os.path.join(ROOT, "mcmc_clv_model_b200", "no_such_lib.so"))
    with pytest.raises(ImportError, match="no CPU fallback"):
        L.load()


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm: the unmodified reference from baseline/_ref when baseline/make_ref.py has
    made that copy, else the oracle port; all host cores) prints exactly one JSON line with the contract's keys."""
    import json
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--ref-customers-per-proc", "4000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "customer_updates_per_sec" and d["higher_is_better"] is True
    have_ref = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "bivariate_mcmc.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 1e4 and "workload" in d["config"]
    if have_ref:
        assert d["cpu_port"]["kind"] == "port" and d["cpu_port"]["value"] > d["value"] * 0.5     # the port is the faster restatement


def test_reference_copy_recipe_is_verbatim():
    """baseline/make_ref.py copies the reference's sampler modules unmodified (sha256 recorded in the manifest)."""
    import hashlib
    ref = "/root/reference/src/models/bivariate/mcmc.py"
    dst = os.path.join(ROOT, "baseline", "_ref", "bivariate_mcmc.py")
    if not (os.path.exists(ref) and os.path.exists(dst)):
        pytest.skip("reference tree or baseline/_ref not present")
    assert hashlib.sha256(open(ref, "rb").read()).hexdigest() == hashlib.sha256(open(dst, "rb").read()).hexdigest()


def test_parallel_host_copy_pool():
    """The host side of the staged transfers (a pool of copy threads splitting one memcpy) needs no device: every
    size and alignment must arrive intact, including concurrent callers (chains-over-GPUs runs one sampler per thread)."""
    import threading
    lib = L.load()
    rng = np.random.default_rng(0)
    src = rng.integers(0, 256, 64 * 1024 * 1024 + 77, dtype=np.uint8)
    for nbytes, so, do in ((0, 0, 0), (1, 3, 5), (4095, 1, 2), (2 << 20, 0, 0), ((2 << 20) + 1, 7, 3), (48 << 20, 13, 1),
                           (64 << 20, 0, 64)):
        dst = np.zeros(nbytes + do + 16, dtype=np.uint8)
        nthr = lib.clv_debug_host_copy(dst[do:].ctypes.data, src[so:].ctypes.data, nbytes)
        assert nthr >= 1
        np.testing.assert_array_equal(dst[do:do + nbytes], src[so:so + nbytes])
        assert not dst[:do].any() and not dst[do + nbytes:].any()
    outs = [np.zeros(24 << 20, dtype=np.uint8) for _ in range(4)]
    def work(i):
        for _ in range(3):
            lib.clv_debug_host_copy(outs[i].ctypes.data, src[i * 1000:].ctypes.data, outs[i].size)
    th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    for i in range(4):
        np.testing.assert_array_equal(outs[i], src[i * 1000:i * 1000 + outs[i].size])


def test_first_touch_threads_leave_the_data_alone():
    """clv_run touches the pages of the caller's level-1 array ahead of the copies (`lock or byte, 0`: the write fault is
    taken, the byte keeps its value).  No device needed: a buffer that already holds data stays intact, also while
    another thread keeps writing to it, and a fresh buffer reads as zeros afterwards."""
    import threading
    lib = L.load()
    rng = np.random.default_rng(1)
    src = rng.integers(0, 256, (48 << 20) + 123, dtype=np.uint8)
    dst = src.copy()
    assert lib.clv_debug_first_touch(dst.ctypes.data, dst.size, 4) == 0
    np.testing.assert_array_equal(dst, src)
    for _ in range(3):                   # a copy delivers data while the touchers run over the same pages: nothing is lost
        dst.fill(0)
        th = threading.Thread(target=lambda: [lib.clv_debug_first_touch(dst.ctypes.data, dst.size, 4) for _ in range(4)])
        th.start()
        np.copyto(dst, src)
        th.join()
        np.testing.assert_array_equal(dst, src)
    fresh = np.empty(64 << 20, dtype=np.uint8)
    assert lib.clv_debug_first_touch(fresh[5:].ctypes.data, fresh.size - 5, 3) == 0
    assert lib.clv_debug_first_touch(None, 1, 1) == -1
