#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py into profiles/ (shares per kernel).
usage: summarize_launches.py LAUNCHES.csv OUT.txt [bench_ms_per_step] [bench_it_per_s]"""
import collections
import csv
import re
import sys

src, out = sys.argv[1], sys.argv[2]
ms_step = float(sys.argv[3]) if len(sys.argv) > 3 else None
it_s = float(sys.argv[4]) if len(sys.argv) > 4 else None
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mv, gs = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
seq = []
for r in rows[hi + 1:]:
    if len(r) > mv:
        try:
            seq.append((r[kn], float(r[mv].replace(",", "")), int(r[gs].strip("()").split(",")[0])))
        except ValueError:
            pass
gmax = max((g for n, _, g in seq if "apply_tiled_kernel" in n), default=0)


def short(n, g):
    if "apply_tiled_kernel" in n:
        m = re.search(r"apply_tiled_kernel<([^>]*)>", n)
        return "apply_tiled_kernel<" + (m.group(1) if m else "") + "> " + (
            "[whole slab]" if g > 0.5 * gmax else "[1/16 sub-slab, pipelined host path]")
    for k in ("offdiag_march_kernel", "k_xr", "k_p", "k_s", "k_dot2", "k_dot1", "k_init", "scale_copy_kernel",
              "build_offmask_kernel", "k_store_hist", "recip_copy_kernel", "fill_kernel", "q_pq", "q_vw", "q_xd", "q_xr",
              "q_eps", "q_init", "q_resid", "q_store_hist"):
        if k in n:
            return k
    return "torch/other: " + n[:40]


agg = collections.OrderedDict()
order = []
for n, v, g in seq:
    agg.setdefault(short(n, g), []).append(v)
tot = sum(sum(v) for v in agg.values())
lines = ["# ncu launch list of: python bench.py --steps 10 --warmup 3 --no-cpu --krylov-iters 5   (1x B200, workload C2)",
         "# ncu --metrics gpu__time_duration.sum --clock-control none -c 400; per-launch times are cold-cache and serialised:",
         "# compare SHARES, not absolutes.  Raw CSV under gpurun_out/ (scratch, not committed).", "",
         f"{'kernel':92s} {'n':>4s} {'total us':>10s} {'avg us':>9s} {'share':>7s}"]
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    lines.append(f"{k:92s} {len(v):4d} {sum(v) / 1e3:10.1f} {sum(v) / len(v) / 1e3:9.1f} {sum(v) / tot:7.1%}")
def dot_flag(k):   # template arguments: CMPFIRST, HAS_OFF, HAS_Q, TX, TY, DOT, REV
    m = re.search(r"<([^>]*)>", k)
    a = [x.strip() for x in m.group(1).split(",")] if m else []
    return a[5] if len(a) > 5 else "0"


whole = [v for k, v in agg.items() if k.startswith("apply_tiled") and "whole" in k and dot_flag(k) == "0"]
corr = agg.get("offdiag_march_kernel", [])
if whole:
    a = sum(whole[0]) / len(whole[0]) / 1e3
    # the correction kernel runs once per apply (whole slab) and once per sub-slab launch; take the larger class
    c = max(corr) / 1e3 if corr else 0.0
    lines += ["", f"one device-resident apply step (the bench's timed step) = apply_tiled_kernel {a:.1f} us ({a / (a + c):.1%}) "
                  f"+ offdiag_march_kernel <= {c:.1f} us ({c / (a + c):.1%})"]
    if ms_step:
        lines.append(f"   bench.py measures {ms_step * 1e3:.1f} us per step with CUDA events (warm, back to back): the dominant "
                     f"kernel's share agrees")
    kk = {k: sum(v) / len(v) / 1e3 for k, v in agg.items() if k in ("k_xr", "k_p", "k_s", "k_dot2", "k_dot1")}
    fused = [v for k, v in agg.items() if k.startswith("apply_tiled") and "whole" in k and dot_flag(k) == "1"]
    a2 = sum(fused[0]) / len(fused[0]) / 1e3 if fused else a
    it = (a + c) + (a2 + c) + sum(kk.values())
    lines.append("one BiCGSTAB iteration = apply %.1f us + apply with fused (t,s),(t,t) epilogue %.1f us + " % (a + c, a2 + c) +
                 ", ".join(f"{k} {v:.1f} us ({v / it:.1%})" for k, v in kk.items()) + f"  -> {it:.0f} us under ncu" +
                 (f"; bench: {1e6 / it_s:.0f} us per iteration" if it_s else ""))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[4:]))
