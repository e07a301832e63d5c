"""ctypes binding of libclv_b200.so (include/clv_b200.h).  No CPU fallback: a missing library or a
missing CUDA device raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CLV_B200_LIB") or os.path.join(_HERE, "libclv_b200.so")

MAX_K = 16
RNG_FAST, RNG_STRICT, RNG_INJECTED = 0, 1, 2
COMPAT_REFERENCE, COMPAT_PAPER = 0, 1
SWEEP_AUTO, SWEEP_STREAM, SWEEP_PERSISTENT = 0, 1, 3

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)


class ClvError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libclv_b200 error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [("model_dim", C.c_int32), ("n_cov", C.c_int32), ("n_chains", C.c_int32),
                ("chain_offset", C.c_int32), ("n_mh_steps", C.c_int32), ("rng_mode", C.c_int32),
                ("compat", C.c_int32), ("sweep_mode", C.c_int32), ("device", C.c_int32),
                ("reserved", C.c_int32), ("n_local", C.c_int64), ("n_global", C.c_int64),
                ("gid_offset", C.c_int64), ("seed", C.c_uint64)]


class InitStats(C.Structure):
    _fields_ = [("lam_init", C.c_double), ("mean_mu_init", C.c_double), ("mean_log_s", C.c_double),
                ("omega2", C.c_double), ("max_abs_x", C.c_double), ("xtx", c_double_p)]


class Injected(C.Structure):
    _fields_ = [(n, c_double_p) for n in ("u_z", "e_tau", "u_tau", "t3_l", "t3_m", "u_acc", "n_eta",
                                          "iw_norm", "iw_chi2", "beta_norm")]


class ForecastConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("ncol", C.c_int32), ("n_draws_total", C.c_int64),
                ("n_customers", C.c_int64), ("gid_offset", C.c_int64), ("draw_offset", C.c_int64),
                ("T_star", C.c_double), ("seed", C.c_uint64), ("simulate_spend", C.c_int32),
                ("reserved", C.c_int32), ("sigma_s", C.c_double)]


class GenerateConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("n_cov", C.c_int32), ("n", C.c_int64), ("gid_offset", C.c_int64),
                ("seed", C.c_uint64), ("T_cal_lo", C.c_double), ("T_cal_hi", C.c_double), ("T_star", C.c_double)]


PROGRESS_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.c_int64)

# every symbol include/clv_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "clv_abi_version": (C.c_int, []),
    "clv_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(Config)]),
    "clv_destroy": (None, [C.c_void_p]),
    "clv_last_error": (C.c_char_p, [C.c_void_p]),
    "clv_set_data": (C.c_int, [C.c_void_p, c_int32_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "clv_set_data_columns": (C.c_int, [C.c_void_p, c_int32_p, c_double_p, c_double_p, C.POINTER(c_double_p), c_double_p]),
    "clv_set_hyper": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_double, c_double_p]),
    "clv_init_state": (C.c_int, [C.c_void_p, C.POINTER(InitStats)]),
    "clv_get_init_stats": (C.c_int, [C.c_void_p, C.POINTER(InitStats), c_double_p]),
    "clv_comm_unique_id": (C.c_int, [C.c_void_p]),
    "clv_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "clv_p2p_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "clv_p2p_connect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "clv_p2p_is_cached": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "clv_run": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, c_double_p, c_double_p, c_double_p,
                          PROGRESS_CB, C.c_void_p, C.c_int64]),
    "clv_run_resident": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, c_double_p, c_double_p,
                                   PROGRESS_CB, C.c_void_p, C.c_int64]),
    "clv_resident_draws": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), c_int64_p]),
    "clv_set_fused_forecast": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_uint64]),
    "clv_fused_forecast_result": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_int64_p]),
    "clv_advance": (C.c_int, [C.c_void_p, C.c_int64, C.c_int]),
    "clv_advance_timed": (C.c_int, [C.c_void_p, C.c_int64, c_double_p]),
    "clv_sweeps_done": (C.c_int64, [C.c_void_p]),
    "clv_kernel_launches": (C.c_int64, [C.c_void_p]),
    "clv_set_sweeps_done": (C.c_int, [C.c_void_p, C.c_int64]),
    "clv_set_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "clv_kernel_time_ms": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_int64_p]),
    "clv_get_state": (C.c_int, [C.c_void_p, C.c_int] + [c_double_p] * 7),
    "clv_set_state": (C.c_int, [C.c_void_p, C.c_int] + [c_double_p] * 5),
    "clv_sweep_injected": (C.c_int, [C.c_void_p, C.POINTER(Injected), C.c_int, c_double_p, c_double_p, c_double_p]),
    "clv_forecast": (C.c_int, [C.POINTER(ForecastConfig), c_double_p, c_double_p, c_int64_p, c_double_p]),
    "clv_forecast_dev": (C.c_int, [C.POINTER(ForecastConfig), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "clv_forecast_injected": (C.c_int, [C.POINTER(ForecastConfig), c_double_p, c_double_p, c_double_p, c_double_p,
                                        C.c_int64, c_int64_p, c_int64_p, c_double_p]),
    "clv_forecast_resident": (C.c_int, [C.c_void_p, C.c_double, C.c_uint64, c_int64_p, c_double_p, c_double_p, c_double_p]),
    "clv_upload_draws": (C.c_int, [C.c_void_p, c_double_p, C.c_int64]),
    "clv_posterior_summary": (C.c_int, [C.c_void_p, C.c_double, c_double_p]),
    "clv_weekly_tracking": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int, C.c_uint64, c_double_p]),
    "clv_generate": (C.c_int, [C.POINTER(GenerateConfig), c_double_p, c_double_p, C.c_int, C.c_int, c_int32_p, c_double_p, c_double_p,
                               c_double_p, c_int32_p, c_double_p, c_double_p, c_double_p]),
    "clv_elog2cbs": (C.c_int, [C.c_int, C.c_int64, c_int64_p, c_int32_p, c_double_p, C.c_int32, C.c_int32, C.c_double, c_int64_p,
                               c_int64_p, c_int32_p, c_double_p, c_double_p, c_double_p, c_double_p, c_int32_p, c_double_p,
                               c_double_p, c_int32_p, c_double_p, c_double_p]),
    "clv_standardize": (C.c_int, [C.c_int, C.c_int64, c_double_p, C.c_double, c_double_p, c_double_p, c_double_p]),
    "clv_recode": (C.c_int, [C.c_int, C.c_int64, c_int32_p, c_double_p, C.c_int, c_double_p]),
    "clv_debug_host_copy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "clv_debug_first_touch": (C.c_int, [C.c_void_p, C.c_int64, C.c_int]),
    "clv_debug_lockstep_advance": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int64]),
    "clv_debug_variates": (C.c_int, [C.c_int, C.c_uint64, C.c_uint32, C.c_int32, C.c_int, C.c_int64, c_double_p, c_double_p, c_double_p]),
    "clv_measure_issue_peaks": (C.c_int, [C.c_int, c_double_p]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises ImportError with build instructions when absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C mcmc_clv_model_b200/csrc`.  mcmc_clv_model_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            if os.environ.get("CLV_B200_LIB"):      # an older build loaded for an A/B comparison (tools/kernel_ab.py)
                continue
            raise
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code, handle=None):
    if code != 0:
        msg = load().clv_last_error(handle)
        raise ClvError(code, msg.decode() if msg else "?")


def dptr(a):
    """double* of a C-contiguous float64 ndarray (None -> NULL)."""
    if a is None:
        return None
    assert a.dtype.name == "float64" and a.flags.c_contiguous, (a.dtype, a.flags)
    return a.ctypes.data_as(c_double_p)
