#!/usr/bin/env python
"""Where does the host side of the level-1 transfer go?  The C2-shaped run (4 chains x 23 570 customers, 12 GB of draws)
into (a) a fresh np.empty array, (b) the same array again (every page already there), (c) a fresh array advised
MADV_HUGEPAGE -- with AnonHugePages of the process before / after."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcmc_clv_model_b200 import Sampler

def anon_huge_kb():
    for ln in open("/proc/self/smaps_rollup"):
        if ln.startswith("AnonHugePages"):
            return int(ln.split()[1])
    return -1

d = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "cdnow_full.npz"))
n = d["x"].size
args = (d["x"].astype(np.int32), d["t_x"], d["T_cal"], [d["first_sales_scaled"].astype(float)])
libc = ctypes.CDLL(None, use_errno=True)
shape = (4, 4000, n, 4)
with Sampler(*args, chains=4, seed=42) as s:
    s.run(100, 50, 1)                                   # warm-up of the call path and the staging ring
    def timed(label, out):
        t0 = time.perf_counter()
        s.run(10000, 4000, 1, out=out)
        dt = time.perf_counter() - t0
        print(f"{label}: {dt:.3f} s wall ({out.nbytes / 1e9:.1f} GB; AnonHugePages {anon_huge_kb() / 1e6:.2f} GB)", flush=True)
    t0 = time.perf_counter(); s.run(10000, 4000, 1, store_level1=False); print(f"no level-1 output: {time.perf_counter() - t0:.3f} s", flush=True)
    a = np.empty(shape)
    timed("fresh np.empty            ", a)
    timed("same array again (touched)", a)
    timed("same array, third time    ", a)
    del a
    b = np.empty(shape)
    lo = (b.ctypes.data + 4095) & ~4095
    hi = (b.ctypes.data + b.nbytes) & ~4095
    rc = libc.madvise(ctypes.c_void_p(lo), ctypes.c_size_t(hi - lo), 14)      # MADV_HUGEPAGE
    print("madvise(MADV_HUGEPAGE) rc", rc, "errno", ctypes.get_errno())
    timed("fresh + MADV_HUGEPAGE     ", b)
    timed("same array again          ", b)
