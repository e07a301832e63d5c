#!/usr/bin/env python
"""Summarise ncu outputs (run here, no GPU needed) into profiles/:
   tools/ncu_summary.py <launches.csv> <prof.ncu-rep> <round tag> <customers per launch>"""
import collections, csv, io, json, os, re, subprocess, sys

launches, rep, tag, ncust = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_md = os.path.join(ROOT, "profiles", f"{tag}_sweep_ncu_summary.md")
out_js = os.path.join(ROOT, "profiles", f"{tag}_sweep_metrics.json")

agg = collections.OrderedDict()
if os.path.exists(launches):
    rows = list(csv.reader(l for l in open(launches) if not l.startswith("==")))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    for r in rows[1:]:
        if len(r) > vi:
            agg.setdefault(re.sub(r"\(.*", "", r[ki]), []).append(float(r[vi].replace(",", "")))
STEP = ("k_sweep", "k_level2", "k_persistent")          # the kernels of one Gibbs sweep
ONCE = ("k_split_columns", "k_init_quantities", "k_init_state", "k_derive_params", "k_stats_only")   # once per data set
tot = sum(sum(v) for k, v in agg.items() if any(t in k for t in STEP)) or 1.0

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
h, units, vals = r[0], r[1], r[2]
get = lambda n: float(vals[h.index(n)].replace(",", "")) if n in h else None
unit = lambda n: units[h.index(n)] if n in h else ""
def to_bytes(n):
    v, u = get(n), unit(n).lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
names = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
         "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
         "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
         "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active",
         "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
m = {n: get(n) for n in names}
dr, dw = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
stalls = {n.replace("smsp__pcsamp_warps_issue_stalled_", ""): get(n) for n in h
          if n.startswith("smsp__pcsamp_warps_issue_stalled_") and not n.endswith("_not_issued")}
js = {"kernel": vals[h.index("Kernel Name")] if "Kernel Name" in h else "k_sweep", "customers_per_launch": ncust,
      "duration_us": m["gpu__time_duration.sum"] * {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3,
                                                     "ms": 1e3, "second": 1e6}.get(unit("gpu__time_duration.sum"), 1.0),
      "dram_bytes_read": dr, "dram_bytes_write": dw, "dram_bytes_per_launch": dr + dw,
      "dram_bytes_per_customer": (dr + dw) / ncust, "warp_instructions": m["smsp__inst_executed.sum"],
      "metrics": m, "stall_samples": stalls}
js["warp_instructions_per_warp_sweep"] = m["smsp__inst_executed.sum"] / (ncust / 32)
json.dump(js, open(out_js, "w"), indent=1)
with open(out_md, "w") as f:
    f.write(f"# {tag}: ncu summary of the sweep kernel (`ncu --set full --clock-control none`, {ncust} customers per launch, B200)\n\n")
    f.write("Source reports: `gpurun_out/` (scratch); this file and the JSON beside it are the committed summaries.\n\n")
    f.write("## Launch list (`--metrics gpu__time_duration.sum`; cold-cache, serialised: compare shares)\n\n| kernel | launches | mean us | total ms | share of step kernels |\n|---|---|---|---|---|\n")
    for k, v in agg.items():
        share = (f"{100 * sum(v) / tot:.1f} %" if any(t in k for t in STEP) else
                 "(once per data set)" if any(t in k for t in ONCE) else "(not part of a step)")
        f.write(f"| `{k}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {sum(v) / 1e6:.3f} | {share} |\n")
    f.write(f"\n## `{js['kernel']}` (full set)\n\n| metric | value |\n|---|---|\n")
    f.write(f"| duration | {js['duration_us']:.1f} us |\n| DRAM read / write per launch | {dr / 1e6:.1f} MB / {dw / 1e6:.1f} MB = {(dr + dw) / ncust:.1f} B per customer (algorithmic 84 B) |\n")
    f.write(f"| warp instructions per warp per sweep | {js['warp_instructions_per_warp_sweep']:.0f} (20 MH steps) |\n")
    for n in names[2:]:
        f.write(f"| {n} | {m[n]} |\n")
    f.write("\n## Warp-state samples (pc sampling)\n\n| reason | samples |\n|---|---|\n")
    for k, v in sorted(stalls.items(), key=lambda t: -(t[1] or 0)):
        f.write(f"| {k} | {v:.0f} |\n")
print('wrote', out_md, out_js)
