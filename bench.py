#!/usr/bin/env python
"""bench.py -- customer-updates/s (and ESS/s) of the Abe (2009) sampler; headline workload C4
(synthetic 10 M customers x 4 covariates, 1 chain, 20 MH steps; BASELINE.json configs[3]).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (torchrun for N > 1)
  python bench.py --impl reference [--steps K] [--warmup W]      # the CPU arm: the UNMODIFIED reference (baseline/_ref,
                                                                 # made by baseline/make_ref.py) on all host cores

A "step" is one Gibbs sweep of every customer.  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import hashlib
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_C4 = 10_000_000
S_MH = 20
K_COV = 5
# algorithmic HBM bytes per customer-update, f64 state, K=5: x(4) + t_x(8) + T_cal(8) + 4 covariates(32)
# read, log lambda / log mu read + written (32)
BYTES_PER_UPDATE = 4 + 8 + 8 + 8 * (K_COV - 1) + 32
METRIC = "customer_updates_per_sec"
WORKLOAD = "C4: synthetic bivariate Pareto/NBD, 10M customers x 4 covariates, 1 chain, 20 MH steps"
GOLD = os.path.join(ROOT, "tests", "golden")
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
SWEEP_PROFILE = ("r02_sweep_metrics.json", "r01_sweep_metrics.json")       # newest first


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.samples, self.t0, self.t1 = index, None, [], None, None

    def start(self):
        """Start polling (nvidia-smi needs ~0.3 s to come up: call well before the timed region)."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)

            def pump():
                for ln in self.proc.stdout:
                    self.samples.append((time.time(), ln))
            self.t = threading.Thread(target=pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def __enter__(self):            # the timed region
        self.t0 = time.time()
        return self

    def __exit__(self, *a):
        self.t1 = time.time()

    def stop(self):
        if self.proc:
            time.sleep(0.1)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons, power = [], 0.0, set(), []
        for ts, ln in self.samples:
            if self.t0 is None or not (self.t0 <= ts <= self.t1 + 0.06):
                continue
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = max(mx, float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


def _pin(a):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t.numpy(), t


# ======================================================================================================
# CPU arms
# ======================================================================================================
def _load_reference(which="bivariate"):
    """The unmodified reference module from baseline/_ref (None when the copy is absent)."""
    path = os.path.join(REF_DIR, f"{which}_mcmc.py")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location(f"_clv_reference_{which}", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod            # dataclasses resolve the module through sys.modules
    spec.loader.exec_module(mod)
    return mod


def _synthetic_c4_frame(n, seed):
    """C4 population law drawn with NumPy: X = [1, U(-1,1)^4], theta = exp(X beta + MVN(0, Gamma)), ..."""
    import pandas as pd
    from mcmc_clv_model_b200.synthetic import C4_BETA, C4_GAMMA, C4_T_CAL
    g = np.random.default_rng(seed)
    X = np.column_stack([np.ones(n), g.uniform(-1, 1, (n, K_COV - 1))])
    th = np.exp(X @ C4_BETA + g.multivariate_normal(np.zeros(2), C4_GAMMA, n))
    tau = g.exponential(1.0 / th[:, 1])
    T = g.uniform(*C4_T_CAL, n)
    Te = np.minimum(tau, T)
    x = g.poisson(th[:, 0] * Te)
    t_x = np.where(x > 0, Te * g.random(n) ** (1.0 / np.maximum(x, 1)), 0.0)
    df = pd.DataFrame({"x": x.astype(np.int64), "t_x": t_x, "T_cal": T})
    for k in range(1, K_COV):
        df[f"cov{k}"] = X[:, k]
    return df, X


def _ref_worker(args):
    """One process of the CPU arm: the reference's own mcmc_draw_parameters (or the oracle port) on one customer shard."""
    n, steps, warmup, seed, kind = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    df, X = _synthetic_c4_frame(n, seed)
    if kind == "reference":
        m = _load_reference("bivariate")
        t0 = time.perf_counter()
        m.mcmc_draw_parameters(df, covariates=[f"cov{k}" for k in range(1, K_COV)], mcmc=steps, burnin=warmup, thin=steps, chains=1,
                               seed=seed, trace=0, n_mh_steps=S_MH)
        return time.perf_counter() - t0, warmup + steps
    from oracle import abe_oracle as ao
    from oracle.streams import NumpyOrderStreams
    cbs = ao.Cbs(x=df["x"].to_numpy(), t_x=df["t_x"].to_numpy(), T_cal=df["T_cal"].to_numpy(), X=np.asfortranarray(X))
    hyper = ao.default_hyper(K_COV, 2)
    st = ao.init_state(cbs, hyper, 2)
    src = NumpyOrderStreams(np.random.default_rng(seed))
    for s in range(warmup):
        src.begin_sweep(1 + s)
        ao.sweep(cbs, st, hyper, src, 2, S_MH)
    t0 = time.perf_counter()
    for s in range(steps):
        src.begin_sweep(1 + warmup + s)
        ao.sweep(cbs, st, hyper, src, 2, S_MH)
    return time.perf_counter() - t0, steps


def _cpu_pool_rate(kind, procs, n_per, steps, warmup):
    import multiprocessing as mp
    with mp.get_context("fork").Pool(procs) as pool:
        t0 = time.perf_counter()
        res = pool.map(_ref_worker, [(n_per, steps, warmup, 1000 + p, kind) for p in range(procs)])
        wall = time.perf_counter() - t0
    t = max(r[0] for r in res)
    sweeps = res[0][1]
    return procs * n_per * sweeps / t, t, sweeps, wall


def run_reference(args):
    """CPU arm: the reference's own sampler, unmodified (baseline/_ref; the oracle port when that copy is absent), on all
    host cores -- one independent shard of the C4 population per process, N/cores customers each (bounded at 250 000
    per process so that the run ends within a few minutes)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    kind = "reference" if _load_reference("bivariate") is not None else "port"
    per_full = args.customers // procs
    n_per = int(min(per_full, args.ref_customers_per_proc))
    value, t, sweeps, wall = _cpu_pool_rate(kind, procs, n_per, args.steps, args.warmup)
    sample = (f"{procs} processes x {n_per} synthetic C4 customers (the 10M workload split over {procs} cores would be {per_full} "
              f"each), {sweeps} sweeps timed per process"
              + (" through the reference's mcmc_draw_parameters(burnin=W, mcmc=K): its API has no separate warm-up, so W+K sweeps "
                 "and the chain initialisation are inside the timed call" if kind == "reference" else " (oracle port; warm-up untimed)"))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "customer-updates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / sweeps * (args.customers / (procs * n_per)),
            "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": sample},
            "cpu_baseline": {"value": value, "unit": "customer-updates/s", "cores": procs, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "customer-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": wall}
    if kind == "reference" and not args.no_port_figure:
        pv, pt, ps, _ = _cpu_pool_rate("port", procs, min(n_per, 50_000), min(args.steps, 5), 1)
        line["cpu_port"] = {"value": pv, "unit": "customer-updates/s", "cores": procs, "kind": "port",
                            "sample": f"{procs} processes x {min(n_per, 50_000)} customers x {ps} sweeps (NumPy oracle port, the second figure)"}
    print(json.dumps(line), flush=True)


def cpu_reference_config_rate(which, data, cov, sweeps, chains=1):
    """ms per sweep-chain of the UNMODIFIED reference on a real-data configuration (bounded sample: `sweeps` sweeps, one
    chain, one core -- the reference runs its chains sequentially on one core, bi:481-485)."""
    import pandas as pd
    m = _load_reference(which)
    if m is None:
        return None
    df = pd.DataFrame({k: data[k] for k in data})
    fn = m.mcmc_draw_parameters if which == "bivariate" else m.mcmc_draw_parameters_rfm_m
    t0 = time.perf_counter()
    fn(df, covariates=cov, mcmc=sweeps, burnin=0, thin=sweeps, chains=chains, seed=42, trace=0, n_mh_steps=S_MH)
    dt = time.perf_counter() - t0
    return dt / (sweeps * chains)


# ======================================================================================================
# digests: the same seed and global customer ids must give the same chain on any number of GPUs
# ======================================================================================================
def _mix64(a):
    a = a.copy()
    a ^= a >> np.uint64(30)
    a *= np.uint64(0xBF58476D1CE4E5B9)
    a ^= a >> np.uint64(27)
    a *= np.uint64(0x94D049BB133111EB)
    a ^= a >> np.uint64(31)
    return a


def level1_hash64(gid0, ll, lm):
    """Order-independent 64-bit hash of a level-1 state: sum over customers (mod 2^64) of a mixed (global id, bits of
    log lambda, bits of log mu).  Sums of shards add up to the hash of the whole."""
    gid = np.arange(gid0, gid0 + ll.size, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = _mix64(gid * np.uint64(0x9E3779B97F4A7C15) + np.uint64(1))
        h = _mix64(h ^ ll.view(np.uint64))
        h = _mix64(h ^ lm.view(np.uint64))
        return int(np.sum(h, dtype=np.uint64))


# ======================================================================================================
# our arm
# ======================================================================================================
def run_ours(args):
    import torch
    import torch.distributed as dist
    from mcmc_clv_model_b200 import Sampler
    from mcmc_clv_model_b200.distributed import broadcast_unique_id, shard_bounds
    from mcmc_clv_model_b200.synthetic import C4_BETA, C4_GAMMA, C4_SEED, C4_T_CAL, generate_cbs_arrays

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the sampler has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_tot = args.customers
    lo, hi = shard_bounds(n_tot, world)[rank]
    n_loc = hi - lo

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_u64_over_ranks(v):
        if world == 1:
            return v & 0xFFFFFFFFFFFFFFFF
        # two 32-bit halves in int64: no overflow for <= 2^31 ranks, exact mod 2^64 afterwards
        t = torch.tensor([v & 0xFFFFFFFF, v >> 32], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        return (int(t[0].item()) + (int(t[1].item()) << 32)) & 0xFFFFFFFFFFFFFFFF

    # ---- synthetic C4 shard (device generator; global ids => identical customers for any GPU count) ----
    cols = generate_cbs_arrays(n_loc, C4_BETA, C4_GAMMA, T_cal=C4_T_CAL, T_star=39.0, seed=C4_SEED, gid_offset=lo,
                               device=local, with_truth=True)
    # the frame as mcmc_draw_parameters hands it over: x, t_x, T_cal and the covariate COLUMNS (the intercept of bi:468-470
    # is implicit: clv_set_data_columns)
    pinned = {k: _pin(cols[k]) for k in ("x", "t_x", "T_cal")}
    pinned_cov = [_pin(cols["X"][:, k]) for k in range(1, K_COV)]

    def make(rng="fast", src=None):
        # CBS columns from host memory -> device; exact init statistics on the device (+ NCCL when sharded)
        src = src or dict({k: v[0] for k, v in pinned.items()}, cov=[v[0] for v in pinned_cov])
        comm = (broadcast_unique_id(Sampler.comm_unique_id), rank, world) if world > 1 else None
        s = Sampler(src["x"], src["t_x"], src["T_cal"], src["cov"], model_dim=2, chains=1,
                    n_mh_steps=S_MH, seed=args.seed, rng=rng, device=local, n_global=n_tot, gid_offset=lo, comm=comm)
        if world > 1 and args.collective == "p2p":
            from mcmc_clv_model_b200.distributed import connect_p2p
            connect_p2p(s)        # level-2 statistics all-reduced inside k_level2 over NVLink peer mailboxes
        return s

    # ---- device-resident throughput ("value"): W warm-up sweeps, then exactly K timed ones from the reference's own
    #      initial state (bi:367-379; SURVEY Q9: the degenerate start), production launches (no per-kernel events) ----
    clk = ClockSampler(local).start()
    s = make()
    s.advance(args.warmup, sync=True)
    launches0 = s.kernel_launches
    barrier()
    with clk:
        ms = s.advance_timed(args.steps)
    barrier()
    clk.stop()
    ms = max_over_ranks(ms)
    launches = s.kernel_launches - launches0
    value = n_tot * args.steps / (ms * 1e-3)
    # digest of the chain after W + K sweeps: must not depend on the number of GPUs
    st = s.get_state(0)
    h64 = sum_u64_over_ranks(level1_hash64(lo, st["log_lambda"], st["log_mu"]))
    digest = {"after_sweeps": args.warmup + args.steps,
              "level_2_sha256": hashlib.sha256(st["beta"].tobytes() + st["Sigma"].tobytes()).hexdigest(),
              "level_1_hash64": f"{h64:016x}", "beta_0": [float(v) for v in st["beta"][0]], "Sigma_00": float(st["Sigma"][0, 0]),
              "what": "sha256 of (beta, Sigma) and the order-independent 64-bit hash (sum over customers of mix(global id, "
                      "log lambda, log mu), all-reduced over ranks) of the state after warmup + steps sweeps"}
    exp_path = os.path.join(ROOT, "profiles", "digest_expected.json")
    if os.path.exists(exp_path):
        key = f"customers={n_tot},seed={args.seed},warmup={args.warmup},steps={args.steps}"
        exp = json.load(open(exp_path)).get(key)
        digest["expected_key"] = key
        digest["matches_committed"] = None if exp is None else (exp["level_2_sha256"] == digest["level_2_sha256"] and
                                                                exp["level_1_hash64"] == digest["level_1_hash64"])
    del st
    # the same K sweeps again with CUDA events around every kernel (plain launches: events between the kernels switch
    # programmatic dependent launch off) -> the dominant kernel's average duration for the roofline
    s.set_timing(True)
    barrier()
    ms_ev = max_over_ranks(s.advance_timed(args.steps))
    sweep_ms, l2_ms, n_timed = s.kernel_time_ms()
    s.set_timing(False)
    k_ms = sweep_ms / max(n_timed, 1)
    # stationary regime: the chain restarted at the generating parameters (accept rates of a converged chain)
    s.set_state(0, log_lambda=np.log(cols["lambda_true"]), log_mu=np.log(cols["mu_true"]), beta=C4_BETA, Sigma=C4_GAMMA)
    s.advance(args.warmup, sync=True)
    barrier()
    ms_st = max_over_ranks(s.advance_timed(args.steps))
    stationary = {"value": n_tot * args.steps / (ms_st * 1e-3), "unit": "customer-updates/s", "ms_per_step": ms_st / args.steps,
                  "what": "chain restarted at the generating parameters (set_state), W warm-up + K timed sweeps"}
    s.close()

    peak, which = _peaks()
    hbm_gbs = BYTES_PER_UPDATE * n_loc / (k_ms * 1e-3) / 1e9
    roofline_hbm = {"bound": "hbm", "achieved": hbm_gbs, "peak": peak, "unit": "GB/s", "frac": hbm_gbs / peak,
                    "traffic": None, "peak_source": which, "kernel": "k_sweep2<2,FAST> (two customers per thread)", "kernel_ms": k_ms,
                    "algorithmic_bytes_per_customer_update": BYTES_PER_UPDATE,
                    "algorithmic_bytes_per_launch": BYTES_PER_UPDATE * n_loc,
                    "note": "secondary roof: the sweep kernel is not HBM bound (see `roofline`)"}
    prof = next((os.path.join(ROOT, "profiles", f) for f in SWEEP_PROFILE if os.path.exists(os.path.join(ROOT, "profiles", f))), None)
    pj = None
    if prof:
        try:
            pj = json.load(open(prof))
            roofline_hbm["traffic"] = pj["dram_bytes_per_customer"] * n_loc
            roofline_hbm["traffic_source"] = f"profiles/{os.path.basename(prof)} (ncu --set full: dram bytes per customer x customers per launch)"
        except Exception:
            pj = None

    # ---- the binding roof: instruction issue ---------------------------------------------------------
    smc = torch.cuda.get_device_properties(local).multi_processor_count
    mhz = clk.summary()["sm_mhz"] or 1965.0
    peak_issue = smc * 4 * mhz * 1e6                       # one warp-instruction per scheduler per clock
    kname = (pj or {}).get("kernel", "k_sweep2<2,FAST>").replace("void ", "").replace("(SweepArgs)", "")
    roofline = {"bound": "issue", "kernel": kname, "kernel_ms": k_ms, "unit": "warp-inst/s", "peak": peak_issue,
                "peak_source": f"{smc} SMs x 4 schedulers x {mhz:.0f} MHz (SM clock sampled in the timed region)",
                "kernel_share_of_step": (sweep_ms / max(ms_ev, 1e-9)) if world == 1 else sweep_ms / max(sweep_ms + l2_ms, 1e-9),
                "level2_kernel_us": 1e3 * l2_ms / max(n_timed, 1), "achieved": None, "frac": None, "traffic": roofline_hbm["traffic"]}
    if pj:
        wi = pj["warp_instructions_per_warp_sweep"]        # ncu: smsp__inst_executed.sum / warps of one launch
        ach = (n_loc / 32.0) * wi / (k_ms * 1e-3)
        roofline.update({"achieved": ach, "frac": ach / peak_issue, "warp_instructions_per_warp_sweep": wi,
                         "algorithmic_units": "warp-instructions per launch = customers/32 x warp-instructions per warp-sweep (ncu) ",
                         "ncu_issue_active_pct": pj["metrics"].get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                         "source": f"instruction count from profiles/{os.path.basename(prof)}, duration from this run's CUDA events"})
    if rank == 0 and not args.no_peaks:
        import ctypes as C
        from mcmc_clv_model_b200 import _lib as L
        pk = (C.c_double * 4)()
        L.check(L.load().clv_measure_issue_peaks(local, pk))
        roofline["pipe_peaks_gops"] = {"ffma": pk[0], "imad": pk[1], "mufu_ex2": pk[2], "dfma": pk[3]}
        roofline["customer_mh_steps_per_s_per_gpu"] = n_loc * S_MH / (k_ms * 1e-3)

    # ---- end to end through the C-ABI with host buffers ----------------------------------------------
    def e2e_once(src):
        barrier()
        t0 = time.perf_counter()
        s2 = make(src=src)
        # one kept level-1 draw (the first sweep) -> a fresh host array, D2H included
        out = s2.run(0, args.steps, args.steps, store_level1=True)
        chk = float(out["level_2"][0, 0, 0]) + float(out["level_1"][0, 0, 0, 0])
        s2.close()
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        assert np.isfinite(chk)
        return dt, out["level_2"].nbytes + out["loglik_sum"].nbytes

    sw = make()                     # untimed warm-up of the same call sequence (first-use costs of the draw buffers / copy path)
    sw.run(0, max(args.warmup, 1), max(args.warmup, 1), store_level1=True)
    sw.close()
    e2e_runs = [e2e_once(None) for _ in range(2)]
    e2e_s, small = min(e2e_runs)
    h2d = n_loc * (4 + 8 + 8 + 8 * (K_COV - 1))
    d2h = n_loc * 32 + small
    e2e = {"value": n_tot * args.steps / e2e_s, "unit": "customer-updates/s", "h2d_bytes_per_step": h2d / args.steps,
           "d2h_bytes_per_step": d2h / args.steps, "seconds": e2e_s, "seconds_all_runs": [r[0] for r in e2e_runs],
           "what": "best of two timed runs of: clv_create + clv_set_data_columns (page-locked host CBS columns -> device) + clv_init_state + clv_run(K sweeps, 1 kept "
                   "level-1 draw -> host) + clv_destroy, i.e. everything mcmc_draw_parameters does after the DataFrame is unpacked"}
    pageable = {k: np.array(cols[k], copy=True) for k in ("x", "t_x", "T_cal")}          # ordinary NumPy arrays, as the API receives them
    pageable["cov"] = [np.array(cols["X"][:, k], copy=True) for k in range(1, K_COV)]
    e2e_p_runs = [e2e_once(pageable)[0] for _ in range(2)]
    e2e_p = min(e2e_p_runs)
    e2e_pageable = {"value": n_tot * args.steps / e2e_p, "unit": "customer-updates/s", "seconds": e2e_p, "seconds_all_runs": e2e_p_runs,
                    "what": "the same call sequence from ordinary (pageable) NumPy columns, as mcmc_draw_parameters receives them"}
    del pageable

    # ---- ESS/s at this GPU count: chains across GPUs (bi:485-501: chains are independent), no communication ----
    ess = None
    if not args.no_ess:
        ess = ess_block(args, world, rank, local, max_over_ranks, barrier)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline (N=1 only): bounded sample of the same workload on the host cores ------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        kind = "reference" if _load_reference("bivariate") is not None else "port"
        n_s, sw_ = (500_000, 5) if kind == "reference" else (500_000, 10)      # 10-25 s of CPU work
        dt, nsw = _ref_worker((n_s, sw_, 0, 1, kind))
        cpu = {"value": n_s * nsw / dt, "unit": "customer-updates/s", "cores": 1, "kind": kind,
               "sample": f"{n_s} synthetic C4 customers x {nsw} sweeps ({dt:.1f} s) on one core: the reference is single-threaded "
                         f"(bi:481-485); `--impl reference` runs it on every core"}

    configs = None
    if world == 1 and not args.no_configs:
        configs = configs_block(local, with_cpu=not args.no_cpu_baseline)

    forecast = None
    if world == 1 and not args.no_forecast:
        forecast = forecast_block(args, local, peak)

    line = {"metric": METRIC, "value": value, "unit": "customer-updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{WORKLOAD}, customer-sharded over {world} GPU(s)",
                       "rng": "Philox4x32-10 counter-based; fp32 SFU proposal variates, fp64 target/accept/state",
                       "l2": "state + data = %.0f MB per GPU %s L2 (126 MB); the same state is re-read every sweep by design"
                             % (n_loc * 68 / 1e6, ">" if n_loc * 68 > 126e6 else "<"),
                       "start": "the reference's own initial state (degenerate Sigma, SURVEY Q9); `stationary` has the converged regime",
                       "customer_mh_steps_per_sec": value * S_MH,
                       "collective": ("none" if world == 1 else args.collective)},
            "gpu_launches": int(launches), "clocks": clk.summary(), "e2e": e2e, "e2e_pageable": e2e_pageable,
            "roofline": roofline, "roofline_hbm": roofline_hbm, "stationary": stationary, "digest": digest,
            "cpu_baseline": cpu, "ess": ess, "configs": configs, "forecast": forecast}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def ess_block(args, world, rank, local, max_over_ranks, barrier):
    """C1 (CDNOW Abe subset, 2 357 customers, K=1, 10 000 + 4 000 sweeps) with 56 chains split over the GPUs of the job
    (strong) and with 56 chains on every GPU (weak; 56 x 19 tiles = 1 064 blocks: the largest multiple of 8 chains whose
    tiles are all co-resident, i.e. that the persistent kernel takes in one round).  ESS = rank-normalised split bulk-ESS over all chains,
    min over the level-2 columns, divided by the wall time of the whole call (create, burn-in, sampling, level-2 to host)."""
    import torch.distributed as dist
    from mcmc_clv_model_b200 import Sampler
    from mcmc_clv_model_b200.diagnostics import min_ess
    d = np.load(os.path.join(GOLD, "cdnow_abe.npz"))
    X1 = np.ones((d["x"].size, 1))
    out = {"config": "C1: CDNOW Abe subset N=2357, K=1, 10000 burn-in + 4000 kept per chain, thin 1, 20 MH steps; chains "
                     "across GPUs (independent chains, no communication)",
           "estimator": "bulk-ESS (rank-normalised, split chains, all chains jointly); Geyer sum beside it",
           "reference_numpy_1core": {"chains": 4, "wall_s": 653.6, "min_ess_geyer": 60, "ess_per_sec": 0.09, "source": "BASELINE.md §2"}}

    def one(total_chains, label):
        per = total_chains // world
        barrier()
        t0 = time.perf_counter()
        with Sampler(d["x"], d["t_x"], d["T_cal"], X1, model_dim=2, chains=per, chain_offset=rank * per, n_mh_steps=S_MH,
                     seed=42, device=local) as s:
            l2 = s.run(10000, 4000, 1, store_level1=False)["level_2"]
        wall = max_over_ranks(time.perf_counter() - t0)
        if world > 1:
            parts = [None] * world if rank == 0 else None
            dist.gather_object(l2, parts, dst=0)
            if rank != 0:
                return
            l2 = np.concatenate(parts, axis=0)
        eb, eg = min_ess(l2, "bulk"), min_ess(l2, "geyer")
        out[label] = {"chains_total": int(l2.shape[0]), "chains_per_gpu": per, "wall_s": wall, "min_ess_bulk": eb, "min_ess_geyer": eg,
                      "ess_per_sec": eb / wall, "ess_geyer_per_sec": eg / wall,
                      "customer_updates_per_sec": 2357 * l2.shape[0] * 14000 / wall}

    one(56, "strong")
    if world > 1:
        one(56 * world, "weak")
    elif rank == 0:
        out["weak"] = out["strong"]
    if rank == 0:
        out["ess_per_sec"] = out["strong"]["ess_per_sec"]
    return out


def configs_block(local, with_cpu=True):
    """The configurations the reference actually runs (BASELINE.json configs[0..2]) through the DROP-IN modules
    (src/models/{bivariate,trivariate}/mcmc.py: DataFrame in, the reference's dict of host arrays out), each beside the
    unmodified reference on one core of the same box (bounded sample, extrapolated)."""
    import pandas as pd
    from mcmc_clv_model_b200.diagnostics import min_ess
    from src.models.bivariate.mcmc import mcmc_draw_parameters
    from src.models.trivariate.mcmc import mcmc_draw_parameters_rfm_m
    abe = dict(np.load(os.path.join(GOLD, "cdnow_abe.npz")))
    full = dict(np.load(os.path.join(GOLD, "cdnow_full.npz")))
    cases = [("C1", "bivariate", abe, [], 4, "run_mcmc_abe.py:61-71: Abe subset N=2357, M1 (K=1), 4 chains x (10000 + 4000)", 100),
             ("C2", "bivariate", full, ["first_sales_scaled"], 2,
              "bivariate/run_mcmc_full.py:137-147 shape: full CDNOW N=23570, K=2, 2 chains x (10000 + 4000), level-1 draws to host", 12),
             ("C3", "trivariate", full, ["gender_F", "age_scaled"], 2,
              "trivariate/run_mcmc_full.py:80-90 shape: full CDNOW N=23570, trivariate K=3, 2 chains x (10000 + 4000)", 12)]
    out = {}
    for name, which, data, cov, chains, what, cpu_sweeps in cases:
        df = pd.DataFrame({k: data[k] for k in data})
        fn = mcmc_draw_parameters if which == "bivariate" else mcmc_draw_parameters_rfm_m
        n = len(df)
        fn(df, covariates=cov, mcmc=40, burnin=40, thin=1, chains=chains, seed=1, trace=0)           # warm-up of the call path
        walls, draws = [], None
        for _ in range(2):              # best of two: 1.2 - 7.5 GB land in fresh NumPy arrays, page-fault bound and noisy (C2: 0.72 - 1.6 s)
            del draws
            t0 = time.perf_counter()
            draws = fn(df, covariates=cov, mcmc=4000, burnin=10000, thin=1, chains=chains, seed=42, trace=0)
            walls.append(time.perf_counter() - t0)
        wall = min(walls)
        l2 = np.asarray(draws["level_2"])
        gb = sum(a.nbytes for a in draws["level_1"]) / 1e9
        e = {"what": what, "wall_s": wall, "wall_s_all_runs": walls, "customer_updates_per_sec": n * chains * 14000 / wall, "level_1_to_host_GB": gb,
             "min_ess_bulk": min_ess(l2, "bulk"), "min_ess_geyer": min_ess(l2, "geyer"),
             "log_likelihood": float(draws["log_likelihood"]), "level_2_mean": [float(v) for v in l2.mean(axis=(0, 1))]}
        e["ess_per_sec"] = e["min_ess_bulk"] / wall
        del draws
        if with_cpu:
            per = cpu_reference_config_rate(which, data, cov, cpu_sweeps)
            if per is not None:
                e["cpu_reference"] = {"kind": "reference", "cores": 1, "seconds_per_sweep_chain": per,
                                      "sample": f"unmodified reference, 1 chain x {cpu_sweeps} sweeps on one core",
                                      "extrapolated_full_run_s": per * 14000 * chains,
                                      "customer_updates_per_sec": n / per}
                e["speedup_vs_cpu_reference_extrapolated"] = per * 14000 * chains / wall
        out[name] = e
    return out


def forecast_block(args, local, peak):
    """C5 (BASELINE.json configs[4]): x* and P(alive) over 1 M synthetic customers x 2 000 posterior draws kept by the sampler
    itself and still resident in HBM (64 GB), T_star = 39; plus the API path with host arrays."""
    from mcmc_clv_model_b200 import Sampler
    from mcmc_clv_model_b200.api import _forecast
    from mcmc_clv_model_b200.synthetic import C4_BETA, C4_GAMMA, C4_SEED, C4_T_CAL, generate_cbs_arrays
    nf, nd = args.forecast_customers, args.forecast_draws
    fc = generate_cbs_arrays(nf, C4_BETA, C4_GAMMA, T_cal=C4_T_CAL, T_star=39.0, seed=C4_SEED + 1, device=local, with_truth=True)
    with Sampler(fc["x"], fc["t_x"], fc["T_cal"], fc["X"], model_dim=2, chains=1, n_mh_steps=S_MH, seed=7, device=local) as sf:
        sf.set_state(0, log_lambda=np.log(fc["lambda_true"]), log_mu=np.log(fc["mu_true"]), beta=C4_BETA, Sigma=C4_GAMMA)
        t0 = time.perf_counter()
        sf.run_resident(20, nd, 1)
        t_sample = time.perf_counter() - t0
        sf.forecast_resident(T_star=39.0, seed=42)                         # warm-up
        best = min(sf.forecast_resident(T_star=39.0, seed=42)["kernel_ms"] for _ in range(3))
        fr = sf.forecast_resident(T_star=39.0, seed=42)
    cells = nf * nd
    gbs = cells * 32 / (best * 1e-3) / 1e9                                  # one 32-byte level-1 row per cell
    forecast = {"config": f"C5: {nf} synthetic customers x {nd} posterior draws (kept by the sampler, resident in HBM: "
                          f"{cells * 32 / 1e9:.0f} GB), T_star=39",
                "cells_per_sec": cells / (best * 1e-3), "kernel_ms": best,
                "kernel": "k_forecast_reduce<4> (every cell's quick path) + k_forecast_deferred<4> (the ~6 % of cells that need more) + k_scale",
                "sampling_s": t_sample,
                "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                             "algorithmic_bytes_per_cell": 32,
                             # ncu --set full (profiles/r02_forecast_ncu_summary.md): 6.444 + 0.508 GB read, 0.016 GB written per 2e8 cells
                             "traffic": cells * (6.443752e9 + 0.508303e9 + 0.0158e9) / 2e8,
                             "traffic_source": "profiles/r02_forecast_ncu_summary.md (dram bytes per cell of both passes x cells)"},
                "mean_x_star": float(fr["mean_x_star"].mean()), "mean_p_alive": float(fr["p_alive"].mean()),
                "holdout_mean_x_star_generated": float(fc["x_star"].mean())}
    # the API path of draw_future_transactions (SURVEY 8d, C5 ii): host (n_draws, N, 4) f64 in, host (n_draws, N)
    # int64 out through clv_forecast -- 40 bytes per cell over PCIe, which is what bounds it
    nd_api = 64
    l1 = np.empty((nd_api, nf, 4))
    l1[:, :, 0] = fc["lambda_true"]; l1[:, :, 1] = fc["mu_true"]; l1[:, :, 2] = fc["tau_true"]
    l1[:, :, 3] = (fc["tau_true"] > fc["T_cal"]).astype(np.float64)
    _forecast(fc["T_cal"], [l1[:4]], 39.0, 42, False, 0.5, devices=[local])      # warm-up (streams, pool)
    t0 = time.perf_counter()
    xs, _ = _forecast(fc["T_cal"], [l1], 39.0, 42, False, 0.5, devices=[local])
    dt = time.perf_counter() - t0
    forecast["api_path"] = {"config": f"{nf} customers x {nd_api} draws, pageable host arrays in and out (the reference's layouts)",
                            "cells_per_sec": nf * nd_api / dt, "seconds": dt, "pcie_GBps": nf * nd_api * 40 / dt / 1e9,
                            "bound": "PCIe + pageable staging (40 B per cell)", "mean_x_star": float(xs.mean())}
    m = _load_reference("bivariate")
    if m is not None and not args.no_cpu_baseline:
        import pandas as pd
        nc, ndc = 100_000, 20
        t0 = time.perf_counter()
        m.draw_future_transactions(pd.DataFrame({"T_cal": fc["T_cal"][:nc]}), {"level_1": [l1[:ndc, :nc]]}, T_star=39.0, seed=42)
        dtc = time.perf_counter() - t0
        forecast["cpu_reference"] = {"kind": "reference", "cores": 1, "cells_per_sec": nc * ndc / dtc,
                                     "sample": f"unmodified reference draw_future_transactions, {nc} customers x {ndc} draws ({dtc:.1f} s)"}
    return forecast


def main():
    # Libraries (NCCL's version banner, torchrun hints) may write to fd 1: park it on stderr and keep the real
    # stdout for the single JSON line.
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--customers", type=int, default=N_C4)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--collective", default="p2p", choices=["p2p", "nccl"],
                    help="customer-sharded runs: all-reduce of the level-2 statistics fused into k_level2 over peer memory, or NCCL")
    ap.add_argument("--no-ess", action="store_true")
    ap.add_argument("--no-forecast", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--no-peaks", action="store_true")
    ap.add_argument("--no-port-figure", action="store_true")
    ap.add_argument("--forecast-customers", type=int, default=1_000_000)
    ap.add_argument("--forecast-draws", type=int, default=2000)
    ap.add_argument("--ref-customers-per-proc", type=int, default=250_000)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
