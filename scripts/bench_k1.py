#!/usr/bin/env python
"""A/B timing of the operator kernel under the tuning environment of the calling shell (FDFD_K1_GEN, FDFD_RP_NWC,
FDFD_RP_NST, FDFD_RP_NCHUNK, ...): GDOF/s and fraction of the HBM roofline on C2 / C3 / C4 / C5-unit, plus a parity
check of the kernel under test against the general (one-thread-per-cell) kernel on the same device.  One JSON line per config."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import workloads
import maxwellfdm_jl_b200 as fb

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
TAG = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("FDFD_"))


def measure(name, w, steps=50, check=True, kry=0):
    objs = "shapes" in w
    A = workloads.make_operator_from_objects(w, device=0) if objs else workloads.make_operator(w, device=0)
    n = A.n
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(n, 2, device="cuda", dtype=torch.float64, generator=g).view(torch.complex128).reshape(-1)
    y = torch.empty_like(x)
    A.bench_apply(x, y, warmup=5, iters=1)
    ms, _ = A.bench_apply(x, y, warmup=0, iters=steps)
    ms /= steps
    off = A.offdiag_fraction if w["full_eps"] else 0.0
    bpd = 32 + A.mass_bytes_per_dof + A.offdiag_bytes_per_dof * off
    out = {"tag": TAG, "config": name, "ms": round(ms, 5), "gdof_s": round(n / ms / 1e6, 2),
           "hbm_frac": round(bpd * n / (ms * 1e-3) / 1e9 / PEAK, 4), "bpd": round(bpd, 2)}
    if kry:
        b = torch.randn(n, 2, device="cuda", dtype=torch.float64, generator=g).view(torch.complex128).reshape(-1)
        xs = torch.zeros_like(b)
        t = A.bench_solve(b, xs, "bicgstab", warmup=2, iters=kry)
        out["bicgstab_it_s"] = round(kry / (t * 1e-3), 1)
    if check:
        B = (workloads.make_operator_from_objects(w, device=0, kernel=fb._lib.KERNEL_NAIVE) if objs
             else workloads.make_operator(w, device=0, kernel=fb._lib.KERNEL_NAIVE))
        y2 = B @ x
        out["rel_vs_general_kernel"] = float((torch.linalg.norm(y - y2) / torch.linalg.norm(y2)).item())
        B.close()
    A.close()
    return out


if __name__ == "__main__":
    torch.cuda.set_device(0)
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["c2", "c3", "c4", "c5"]
    check = "--no-check" not in sys.argv
    kry = 20 if "--krylov" in sys.argv else 0
    if "c2" in which:
        print(json.dumps(measure("C2", workloads.c2_waveguide(), check=check, kry=kry)), flush=True)
    if "c2d" in which:
        import numpy as np
        w = workloads.c2_waveguide()
        rng = np.random.default_rng(7)
        for (v, u) in ((0, 1), (0, 2), (1, 2)):
            pert = 0.05 * (rng.random(w["eps"].shape[:3]) - 0.5)
            w["eps"][..., v, u] = pert
            w["eps"][..., u, v] = pert
        print(json.dumps(measure("C2 dense", w, check=check, kry=kry)), flush=True)
    if "c3" in which:
        print(json.dumps(measure("C3", workloads.c3_phc_slab(), check=check, kry=kry)), flush=True)
    if "c4" in which:
        print(json.dumps(measure("C4 obj", workloads.c4_objects(), steps=20, check=check, kry=min(kry, 5))), flush=True)
    if "c5" in which:
        print(json.dumps(measure("C5 obj", workloads.c5_objects(), steps=20, check=check, kry=min(kry, 5))), flush=True)
