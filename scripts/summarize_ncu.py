#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small text file for profiles/ (read on the CPU box).
usage: summarize_ncu.py REPORT.ncu-rep OUT.txt [algorithmic_bytes_per_launch]"""
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
alg = float(sys.argv[3]) if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
lines = [f"# ncu summary of {rep}", "# produced by scripts/summarize_ncu.py (ncu --set full --clock-control none); cold-cache, serialised launches", ""]
for r in rows[2:]:
    g = lambda k: (r[h.index(k)], units[h.index(k)]) if k in h else ("n/a", "")
    lines.append(f"## launch {r[h.index('ID')]}: {r[h.index('Kernel Name')][:110]}")
    for k in KEYS:
        v, u = g(k)
        lines.append(f"  {k:72s} {v:>16s} {u}")
    try:
        def tobytes(k):
            v, u = g(k)
            m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            return float(v.replace(",", "")) * m
        tr = tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")
        dv, du = g("gpu__time_duration.sum")
        dur = float(dv.replace(",", "")) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}[du]
        lines.append(f"  -> DRAM traffic {tr/1e6:.1f} MB per launch, {tr/dur/1e9:.0f} GB/s under ncu")
        if alg:
            lines.append(f"  -> algorithmic bytes {alg/1e6:.1f} MB per launch: traffic/algorithmic = {tr/alg:.3f}; "
                         f"algorithmic GB/s under ncu = {alg/dur/1e9:.0f}")
    except Exception as e:  # noqa: BLE001
        lines.append(f"  (traffic summary failed: {e})")
    st = [(h[i], float(r[i].replace(",", ""))) for i in range(len(h))
          if "pcsamp_warps_issue_stalled" in h[i] and "not_issued" not in h[i] and r[i] not in ("", "n/a")]
    tot = sum(v for _, v in st) or 1
    lines.append("  warp stall samples: " + ", ".join(
        f"{k.replace('smsp__pcsamp_warps_issue_stalled_', '')}={v / tot:.1%}" for k, v in sorted(st, key=lambda kv: -kv[1])[:10]))
    lines.append("")
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(sass.splitlines()))
if len(rows) > 2:
    hh = rows[1]
    isrc, ins = hh.index("Source"), hh.index("# Samples")
    data = [(int(x[ins] or 0), x[isrc]) for x in rows[2:] if len(x) >= len(hh) and x[ins].isdigit()]
    tot = sum(d[0] for d in data) or 1
    lines.append("## hottest SASS instructions of the first kernel (share of stall samples)")
    for i, d in sorted(enumerate(data), key=lambda t: -t[1][0])[:12]:
        lines.append(f"  #{i:4d} {d[0] / tot:6.1%}  {d[1][:100]}")
    ops = {}
    for _, src in data:
        op = src.split()[0] if not src.startswith("@") else src.split()[1]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + 1
    lines.append("  static SASS mix: " + ", ".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:14]))
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out)
