#!/usr/bin/env python
"""Compact per-launch table from an .ncu-rep (`ncu -i rep --page raw --csv`): the metrics DESIGN.md / bench.py quote."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h, u = rows[hdr], rows[hdr + 1]
want = [("dur_us", "gpu__time_duration.sum"), ("sm_clk_mhz", "smsp__cycles_elapsed.avg.per_second"),
        ("tensor_pipe_%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("sm_busy_%", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("dram_%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("dram_rd_MB", "dram__bytes_read.sum"), ("dram_wr_MB", "dram__bytes_write.sum"),
        ("l2_%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"), ("l2_hit_%", "lts__t_sector_hit_rate.pct"),
        ("regs", "launch__registers_per_thread"), ("smem_KB", "launch__shared_mem_per_block_dynamic"),
        ("grid", "launch__grid_size"), ("warps_active_%", "sm__warps_active.avg.pct_of_peak_sustained_active")]
idx = {}
for k, name in want:
    for i, n in enumerate(h):
        if n == name:
            idx[k] = i
            break


def conv(k, i, v):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    unit = u[i].lower()
    if k == "dur_us":
        x = x / 1e3 if unit.startswith("ns") else (x * 1e3 if unit.startswith("ms") else x)
    if k.endswith("_MB"):
        x = {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(unit, 1.0) * x
    if k == "smem_KB":
        x = {"byte": 1 / 1024, "kbyte": 1.0, "mbyte": 1024.0}.get(unit, 1.0) * x
    if k == "sm_clk_mhz":
        x = {"hz": 1e-6, "khz": 1e-3, "mhz": 1.0, "ghz": 1e3}.get(unit, 1.0) * x
    return f"{x:.1f}" if abs(x) < 1e5 else f"{x:.3g}"


print("| kernel | " + " | ".join(k for k, _ in want if k in idx) + " |")
print("|---|" + "---|" * len(idx))
for v in rows[hdr + 2:]:
    if len(v) != len(h):
        continue
    name = v[h.index("Kernel Name")]
    name = name.replace("void ", "").split("(")[0][-60:]
    print(f"| `{name}` | " + " | ".join(conv(k, idx[k], v[idx[k]]) for k, _ in want if k in idx) + " |")
