#!/usr/bin/env python
"""Summarise the ncu --set full captures of the forecast kernels into profiles/<tag>_forecast_ncu_summary.md:
   tools/ncu_forecast_summary.py <tag> <cells per launch> <name>=<rep> [<name>=<rep> ...]"""
import csv, io, os, subprocess, sys

tag, cells = sys.argv[1], float(sys.argv[2])
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
names = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
         "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
         "launch__registers_per_thread", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
         "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
         "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]
out = os.path.join(ROOT, "profiles", f"{tag}_forecast_ncu_summary.md")
with open(out, "w") as f:
    f.write(f"# {tag}: ncu summary of the forecast kernels (`ncu --set full --clock-control none`, {cells:.3g} cells per launch = "
            f"{cells * 32 / 1e9:.1f} GB of level-1 rows, B200)\n\nSource reports: `gpurun_out/` (scratch).  Per-launch times under ncu are cold-cache "
            "and serialised; the bench numbers (CUDA events, warm) are in `profiles/r02_kernel_ab.txt`.\n")
    for spec in sys.argv[3:]:
        label, rep = spec.split("=", 1)
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        r = list(csv.reader(io.StringIO(raw)))
        h, units = r[0], r[1]
        seen = set()
        for row in r[2:]:
            kn = row[h.index("Kernel Name")].split("(")[0]
            if kn in seen:
                continue
            seen.add(kn)
            get = lambda n: row[h.index(n)] if n in h else ""
            f.write(f"\n## {label}: `{kn}`\n\n| metric | value |\n|---|---|\n")
            for n in names:
                f.write(f"| {n} | {get(n)} {units[h.index(n)] if n in h else ''} |\n")
            try:
                dur = float(get("gpu__time_duration.sum")) * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "msecond": 1e-3, "usecond": 1e-6}.get(units[h.index("gpu__time_duration.sum")], 1e-6)
                f.write(f"| warp instructions per cell | {float(get('smsp__inst_executed.sum')) / cells:.2f} |\n")
                f.write(f"| algorithmic GB/s of this launch (32 B per cell of the whole forecast) | {cells * 32 / dur / 1e9:.0f} |\n")
            except Exception:
                pass
            st = {n.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(row[h.index(n)] or 0) for n in h
                  if n.startswith("smsp__pcsamp_warps_issue_stalled_") and not n.endswith("_not_issued")}
            f.write("\nwarp states (pc samples): " + ", ".join(f"{k} {v:.0f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]) + "\n")
print("wrote", out)
