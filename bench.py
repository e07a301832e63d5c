#!/usr/bin/env python
"""bench.py -- customer-updates/s of the Abe (2009) bivariate sampler on config C4
(synthetic 10 M customers x 4 covariates, 1 chain, 20 MH steps; BASELINE.json configs[3]).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (torchrun for N > 1)
  python bench.py --impl reference [--steps K] [--warmup W]      # the CPU arm (oracle port, all host cores)

A "step" is one Gibbs sweep of every customer.  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
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


def cpu_port_rate(cols, n_sample, sweeps, seed=1):
    """The oracle port (NumPy, the reference's algorithm and RNG call order) on the first n_sample customers:
    customer-updates/s on ONE core (the reference is single-threaded, bi:481-485)."""
    from oracle import abe_oracle as ao
    from oracle.streams import NumpyOrderStreams
    n = min(n_sample, cols["x"].size)
    cbs = ao.Cbs(x=cols["x"][:n].astype(np.int64), t_x=cols["t_x"][:n], T_cal=cols["T_cal"][:n],
                 X=np.asfortranarray(cols["X"][:n]))
    hyper = ao.default_hyper(cbs.K, 2)
    st = ao.init_state(cbs, hyper, 2)
    src = NumpyOrderStreams(np.random.default_rng(seed))
    src.begin_sweep(1)
    ao.sweep(cbs, st, hyper, src, 2, S_MH)                     # warm-up sweep (page-in, BLAS init)
    t0 = time.perf_counter()
    for s in range(sweeps):
        src.begin_sweep(2 + s)
        ao.sweep(cbs, st, hyper, src, 2, S_MH)
    dt = time.perf_counter() - t0
    return n * sweeps / dt, n, dt


def _ref_worker(args):
    n, steps, warmup, seed = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import abe_oracle as ao
    from oracle.streams import NumpyOrderStreams
    g = np.random.default_rng(seed)
    # same population law as the device generator: X = [1, U(-1,1)^4], theta = exp(X beta + MVN(0, Gamma))
    from mcmc_clv_model_b200.synthetic import C4_BETA, C4_GAMMA, C4_T_CAL
    X = np.column_stack([np.ones(n), g.uniform(-1, 1, (n, K_COV - 1))])
    th = np.exp(X @ C4_BETA + g.multivariate_normal(np.zeros(2), C4_GAMMA, n))
    tau = g.exponential(1.0 / th[:, 1])
    T = g.uniform(*C4_T_CAL, n)
    Te = np.minimum(tau, T)
    x = g.poisson(th[:, 0] * Te)
    t_x = np.where(x > 0, Te * g.random(n) ** (1.0 / np.maximum(x, 1)), 0.0)
    cbs = ao.Cbs(x=x.astype(np.int64), t_x=t_x, T_cal=T, X=np.asfortranarray(X))
    hyper = ao.default_hyper(K_COV, 2)
    st = ao.init_state(cbs, hyper, 2)
    src = NumpyOrderStreams(g)
    for s in range(warmup):
        src.begin_sweep(1 + s)
        ao.sweep(cbs, st, hyper, src, 2, S_MH)
    t0 = time.perf_counter()
    for s in range(steps):
        src.begin_sweep(1 + warmup + s)
        ao.sweep(cbs, st, hyper, src, 2, S_MH)
    return time.perf_counter() - t0


def run_reference(args):
    """CPU arm: the reference's algorithm (oracle port; the reference is pure Python/NumPy and is not shipped to
    the GPU box) on all host cores, one independent customer shard per process."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    n_per = 20_000
    with mp.get_context("fork").Pool(procs) as pool:
        t0 = time.perf_counter()
        times = pool.map(_ref_worker, [(n_per, args.steps, args.warmup, 1000 + p) for p in range(procs)])
        wall = time.perf_counter() - t0
    t = max(times)
    value = procs * n_per * args.steps / t
    sample = f"{procs} processes x {n_per} synthetic C4 customers x {args.steps} sweeps (bounded sample of the 10M workload)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "customer-updates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C4: synthetic bivariate Pareto/NBD, 10M customers x 4 covariates, 1 chain, 20 MH steps",
                       "sample": sample},
            "cpu_baseline": {"value": value, "unit": "customer-updates/s", "cores": procs, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "customer-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": wall}
    print(json.dumps(line), flush=True)


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

    # ---- synthetic C4 shard (device generator; global ids => identical customers for any GPU count) ----
    cols = generate_cbs_arrays(n_loc, C4_BETA, C4_GAMMA, T_cal=C4_T_CAL, T_star=39.0, seed=C4_SEED, gid_offset=lo,
                               device=local, with_truth=False)
    pinned = {k: _pin(cols[k]) for k in ("x", "t_x", "T_cal", "X")}

    def make(rng="fast"):
        # CBS columns from pinned host memory -> device; exact init statistics on the device (+ NCCL when sharded)
        comm = (broadcast_unique_id(Sampler.comm_unique_id), rank, world) if world > 1 else None
        s = Sampler(pinned["x"][0], pinned["t_x"][0], pinned["T_cal"][0], pinned["X"][0], model_dim=2, chains=1,
                    n_mh_steps=S_MH, seed=args.seed, rng=rng, device=local, n_global=n_tot, gid_offset=lo, comm=comm)
        if world > 1 and args.collective == "p2p":
            from mcmc_clv_model_b200.distributed import connect_p2p
            connect_p2p(s)        # level-2 statistics all-reduced inside k_level2 over NVLink peer mailboxes
        return s

    # ---- device-resident throughput ("value") -----------------------------------------------------
    clk = ClockSampler(local).start()
    s = make()
    s.advance(args.warmup, sync=True)
    s.set_timing(True)
    launches0 = s.kernel_launches
    barrier()
    with clk:
        ms = s.advance_timed(args.steps)
    barrier()
    clk.stop()
    ms = max_over_ranks(ms)
    launches = s.kernel_launches - launches0
    sweep_ms, l2_ms, n_timed = s.kernel_time_ms()
    s.set_timing(False)
    value = n_tot * args.steps / (ms * 1e-3)
    peak, which = _peaks()
    k_ms = sweep_ms / max(n_timed, 1)
    achieved = BYTES_PER_UPDATE * n_loc / (k_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": which, "kernel": "k_sweep<2,FAST>", "kernel_ms": k_ms,
                "kernel_share_of_step": sweep_ms / (ms if world == 1 else max(sweep_ms + l2_ms, 1e-9)),
                "algorithmic_bytes_per_customer_update": BYTES_PER_UPDATE, "algorithmic_bytes_per_launch": BYTES_PER_UPDATE * n_loc,
                "note": "the sweep is instruction-issue bound (Philox INT32 + FP64 target + MUFU), not HBM bound; see issue_roofline"}
    prof = os.path.join(ROOT, "profiles", "r01_sweep_metrics.json")
    if os.path.exists(prof):
        try:
            # ncu --set full capture of this kernel (profiles/r01_sweep_ncu_summary.md), scaled to this launch's customers
            roofline["traffic"] = json.load(open(prof)).get("dram_bytes_per_customer") * n_loc
            roofline["traffic_source"] = "profiles/r01_sweep_metrics.json (dram bytes per customer x customers per launch)"
        except Exception:
            pass
    s.close()

    # ---- issue-rate roofline (the binding one): measured pipe peaks vs achieved op rates -------------
    issue = None
    if rank == 0:
        import ctypes as C
        from mcmc_clv_model_b200 import _lib as L
        pk = (C.c_double * 4)()
        L.check(L.load().clv_measure_issue_peaks(local, pk))
        issue = {"peaks_gops": {"ffma": pk[0], "imad": pk[1], "mufu_ex2": pk[2], "dfma": pk[3]},
                 "customer_mh_steps_per_s_per_gpu": n_loc * S_MH / (k_ms * 1e-3)}
        if os.path.exists(prof):
            try:
                pj = json.load(open(prof))
                wi = pj["warp_instructions_per_warp_sweep"]                   # ncu: smsp__inst_executed.sum / warps
                smc = torch.cuda.get_device_properties(local).multi_processor_count
                mhz = clk.summary()["sm_mhz"] or 1965.0
                achieved = (n_loc / 32.0) * wi / (k_ms * 1e-3)                  # warp-instructions issued per second
                peak_issue = smc * 4 * mhz * 1e6                                # one warp-instruction per scheduler per clock
                issue.update({"bound": "instruction issue", "warp_instructions_per_warp_sweep": wi,
                              "achieved_warp_inst_per_s": achieved, "peak_warp_inst_per_s": peak_issue,
                              "frac": achieved / peak_issue,
                              "ncu_issue_active_pct": pj["metrics"].get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                              "source": "instruction count per warp from profiles/r01_sweep_metrics.json (ncu), rate from this run's CUDA events"})
            except Exception:
                pass

    # ---- end to end through the C-ABI with host buffers ----------------------------------------------
    # untimed warm-up of the same call sequence (first-use costs of the draw buffers / copy path), then the timed one
    sw = make()
    sw.run(0, max(args.warmup, 1), max(args.warmup, 1), store_level1=True)
    sw.close()
    barrier()
    t0 = time.perf_counter()
    s2 = make()
    # one kept level-1 draw (the first sweep) -> a fresh host array, D2H included (a page-locked destination allocated
    # inside the timed region was measured slower: cudaHostAlloc of 320 MB costs more than the blocking copy)
    out = s2.run(0, args.steps, args.steps, store_level1=True)
    chk = float(out["level_2"][0, 0, 0]) + float(out["level_1"][0, 0, 0, 0])
    s2.close()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    h2d = n_loc * (4 + 8 + 8 + 8 * K_COV)
    d2h = n_loc * 32 + out["level_2"].nbytes + out["loglik_sum"].nbytes
    e2e = {"value": n_tot * args.steps / e2e_s, "unit": "customer-updates/s", "h2d_bytes_per_step": h2d / args.steps,
           "d2h_bytes_per_step": d2h / args.steps, "seconds": e2e_s,
           "what": "clv_create + clv_set_data (pinned host CBS -> device) + clv_init_state + clv_run(K sweeps, 1 kept "
                   "level-1 draw -> host) + clv_destroy, i.e. everything mcmc_draw_parameters does after the DataFrame is unpacked"}
    assert np.isfinite(chk)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline (N=1 only): the oracle port on a bounded sample, one core -----------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, n_s, dt = cpu_port_rate(cols, 200_000, 3)
        cpu = {"value": rate, "unit": "customer-updates/s", "cores": 1, "kind": "port",
               "sample": f"first {n_s} C4 customers x 3 sweeps ({dt:.1f} s), NumPy oracle port, reference RNG order"}

    # ---- ESS/s on C1 (the reference's own run: CDNOW Abe subset, 4 chains x (10000 + 4000) sweeps) -----
    ess = None
    if world == 1 and not args.no_ess:
        from mcmc_clv_model_b200.diagnostics import min_ess
        d = np.load(os.path.join(ROOT, "tests", "golden", "cdnow_abe.npz"))
        X1 = np.ones((d["x"].size, 1))
        t0 = time.perf_counter()
        with Sampler(d["x"], d["t_x"], d["T_cal"], X1, model_dim=2, chains=4, n_mh_steps=20, seed=42, device=local) as s3:
            o = s3.run(10000, 4000, 1, store_level1=True)
        wall = time.perf_counter() - t0
        me_b, me_g = min_ess(o["level_2"], "bulk"), min_ess(o["level_2"], "geyer")
        ess = {"config": "C1: CDNOW Abe subset N=2357, K=1, 4 chains x (10000 burn-in + 4000 kept), thin 1, 20 MH steps",
               "wall_s": wall, "customer_updates_per_sec": 2357 * 4 * 14000 / wall, "min_ess_bulk": me_b,
               "min_ess_geyer": me_g, "ess_per_sec": me_g / wall,
               "reference_numpy_1core": {"wall_s": 653.6, "min_ess_geyer": 60, "ess_per_sec": 0.09, "source": "BASELINE.md §2"}}
        if not args.no_cpu_baseline:
            # the oracle port on the same C1 data, one chain x 200 sweeps on one host core of this box (the reference
            # itself needs 653.6 s for the full run: BASELINE.md)
            from oracle import abe_oracle as ao
            from oracle.streams import NumpyOrderStreams
            cb = ao.Cbs(x=d["x"].astype(np.int64), t_x=d["t_x"].astype(float), T_cal=d["T_cal"].astype(float), X=np.asfortranarray(X1))
            t0 = time.perf_counter()
            ao.run_chain(cb, ao.default_hyper(1, 2), NumpyOrderStreams(np.random.default_rng(42)), mcmc=100, burnin=100, thin=1, D=2)
            dtc = time.perf_counter() - t0
            ess["cpu_port_same_box"] = {"customer_updates_per_sec": 2357 * 200 / dtc, "cores": 1,
                                        "sample": f"C1 data, 1 chain x 200 sweeps ({dtc:.1f} s)",
                                        "extrapolated_full_run_s": dtc / 200 * 14000 * 4}
        # the same data with the GPU filled: 64 chains (the per-sweep latency barely changes, ESS adds up over chains)
        t0 = time.perf_counter()
        with Sampler(d["x"], d["t_x"], d["T_cal"], X1, model_dim=2, chains=64, n_mh_steps=20, seed=42, device=local) as s4:
            o64 = s4.run(10000, 4000, 1, store_level1=False)
        wall64 = time.perf_counter() - t0
        g64 = min_ess(o64["level_2"], "geyer")
        ess["wide"] = {"chains": 64, "wall_s": wall64, "customer_updates_per_sec": 2357 * 64 * 14000 / wall64,
                       "min_ess_geyer": g64, "min_ess_bulk": min_ess(o64["level_2"], "bulk"), "ess_per_sec": g64 / wall64,
                       "level_1": "not stored (level-2 draws only)"}

    # ---- forecast (C5-shaped): x*, P(alive) over customers x posterior draws, draws resident in HBM ----------------
    forecast = None
    if world == 1 and not args.no_forecast:
        nf, nd = args.forecast_customers, args.forecast_draws
        fc = generate_cbs_arrays(nf, C4_BETA, C4_GAMMA, T_cal=C4_T_CAL, T_star=39.0, seed=C4_SEED + 1, device=local, with_truth=True)
        with Sampler(fc["x"], fc["t_x"], fc["T_cal"], fc["X"], model_dim=2, chains=1, n_mh_steps=S_MH, seed=7, device=local) as sf:
            sf.set_state(0, log_lambda=np.log(fc["lambda_true"]), log_mu=np.log(fc["mu_true"]), beta=C4_BETA, Sigma=C4_GAMMA)
            sf.run_resident(20, nd, 1)
            sf.forecast_resident(T_star=39.0, seed=42)                         # warm-up
            best = min(sf.forecast_resident(T_star=39.0, seed=42)["kernel_ms"] for _ in range(3))
            fr = sf.forecast_resident(T_star=39.0, seed=42)
        cells = nf * nd
        gbs = cells * 32 / (best * 1e-3) / 1e9                                  # one 32-byte level-1 row per cell
        forecast = {"config": f"{nf} synthetic customers x {nd} posterior draws (kept by the sampler, resident in HBM), T_star=39",
                    "cells_per_sec": cells / (best * 1e-3), "kernel_ms": best, "kernel": "k_forecast_reduce<4>",
                    "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                 "algorithmic_bytes_per_cell": 32},
                    "mean_x_star": float(fr["mean_x_star"].mean()), "mean_p_alive": float(fr["p_alive"].mean()),
                    "holdout_mean_x_star_generated": float(fc["x_star"].mean())}
        # the API path of draw_future_transactions (SURVEY 8d, C5 ii): host (n_draws, N, 4) f64 in, host (n_draws, N)
        # int64 out through clv_forecast -- 40 bytes per cell over PCIe, which is what bounds it
        from mcmc_clv_model_b200.api import _forecast
        nd_api = 64
        l1 = np.empty((nd_api, nf, 4))
        l1[:, :, 0] = fc["lambda_true"]; l1[:, :, 1] = fc["mu_true"]; l1[:, :, 2] = fc["tau_true"]
        l1[:, :, 3] = (fc["tau_true"] > fc["T_cal"]).astype(np.float64)
        _forecast(fc["T_cal"], [l1[:4]], 39.0, 42, False, 0.5, device=local)      # warm-up (streams, pool)
        t0 = time.perf_counter()
        xs, _ = _forecast(fc["T_cal"], [l1], 39.0, 42, False, 0.5, device=local)
        dt = time.perf_counter() - t0
        forecast["api_path"] = {"config": f"{nf} customers x {nd_api} draws, pageable host arrays in and out (the reference's layouts)",
                                "cells_per_sec": nf * nd_api / dt, "seconds": dt, "pcie_GBps": nf * nd_api * 40 / dt / 1e9,
                                "bound": "PCIe + pageable staging (40 B per cell)", "mean_x_star": float(xs.mean())}
        del l1, xs

    line = {"metric": METRIC, "value": value, "unit": "customer-updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C4: synthetic bivariate Pareto/NBD, {n_tot} customers x 4 covariates, 1 chain, 20 MH steps, "
                                   f"customer-sharded over {world} GPU(s)",
                       "rng": "Philox4x32-10 counter-based; fp32 SFU proposal variates, fp64 target/accept/state",
                       "l2": "state + data = %.0f MB per GPU %s L2 (126 MB); the same state is re-read every sweep by design"
                             % (n_loc * 68 / 1e6, ">" if n_loc * 68 > 126e6 else "<"),
                       "customer_mh_steps_per_sec": value * S_MH,
                       "collective": ("none" if world == 1 else args.collective)},
            "gpu_launches": int(launches), "clocks": clk.summary(), "e2e": e2e, "roofline": roofline,
            "issue_roofline": issue, "cpu_baseline": cpu, "ess": ess, "forecast": forecast}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument("--forecast-customers", type=int, default=1_000_000)
    ap.add_argument("--forecast-draws", type=int, default=500)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
