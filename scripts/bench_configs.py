#!/usr/bin/env python
"""Per-configuration measurements for BASELINE.md section 5 (1 GPU): operator GDOF/s, % of HBM roofline,
BiCGSTAB / QMR iterations/s on the BASELINE.json configurations C1-C3 (+ C2 variants), and the C1 solve
(iterations to 1e-8, true residual, agreement of the two Krylov methods).  Prints one JSON object per line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import workloads
import maxwellfdm_jl_b200 as fb

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def measure(name, w, steps=100, kry=20):
    t_setup = time.perf_counter()
    A = workloads.make_operator_from_objects(w, device=0) if "shapes" in w else workloads.make_operator(w, device=0)
    n = A.n
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(n, 2, device="cuda", dtype=torch.float64, generator=g).view(torch.complex128).reshape(-1)
    y = torch.empty_like(x)
    A.bench_apply(x, y, warmup=5, iters=1)
    t_setup = time.perf_counter() - t_setup       # operator set-up incl. the first applies (objects: rasterisation on the device)
    ms, _ = A.bench_apply(x, y, warmup=0, iters=steps)
    ms /= steps
    off = A.offdiag_fraction if w["full_eps"] else 0.0
    bpd = 32 + A.mass_bytes_per_dof + A.offdiag_bytes_per_dof * off
    b = torch.randn(n, 2, device="cuda", dtype=torch.float64, generator=g).view(torch.complex128).reshape(-1)
    out = {"config": name, "grid": list(w["N"]), "dof": n, "ms_per_apply": ms, "gdof_s": n / ms / 1e6,
           "bytes_per_dof": bpd, "hbm_frac": bpd * n / (ms * 1e-3) / 1e9 / PEAK, "offdiag_block_fraction": off,
           "offdiag_symmetric": bool(w["full_eps"] and A.offdiag_symmetric), "setup_s": t_setup,
           "corr_skip_zero": bool(os.environ.get("FDFD_CORR_SKIP_ZERO"))}
    for method in ("bicgstab", "qmr"):
        xs = torch.zeros_like(b)
        t = A.bench_solve(b, xs, method, warmup=2, iters=kry)
        out[f"{method}_it_s"] = kry / (t * 1e-3)
    A.close()
    return out


def c1_solve():
    w = workloads.c1_vacuum_box()
    A = workloads.make_operator(w, device=0)
    je, jm = workloads.c1_rhs(w)
    b = A.create_b(je)
    res = {"config": "C1 solve (z dipole, rtol 1e-8)"}
    sols = {}
    for method in ("bicgstab", "qmr"):
        t0 = time.perf_counter()
        x, info = A.solve(torch.from_numpy(b).cuda(), method=method, rtol=1e-8, maxit=40000, check_every=50)
        dt = time.perf_counter() - t0
        xs = x.cpu().numpy()
        true = float(np.linalg.norm(A @ xs - b) / np.linalg.norm(b))
        sols[method] = xs
        res[method] = {"iters": info["iters"], "converged": info["converged"], "relres": info["relres"],
                       "true_relres": true, "seconds": dt, "it_s": info["iters"] / dt}
    res["field_agreement_bicgstab_vs_qmr"] = float(np.linalg.norm(sols["bicgstab"] - sols["qmr"]) /
                                                   np.linalg.norm(sols["qmr"]))
    A.close()
    return res


if __name__ == "__main__":
    torch.cuda.set_device(0)
    if "--skip-small" not in sys.argv:
        print(json.dumps(measure("C1 vacuum box 40^3 + PML", workloads.c1_vacuum_box(), steps=300, kry=50)))
        print(json.dumps(measure("C2 Si waveguide 200^3, full eps (sparse off-diagonals)", workloads.c2_waveguide())))
        print(json.dumps(measure("C3 PhC slab 256x256x128, Bloch x/y, PML z", workloads.c3_phc_slab())))
        print(json.dumps(c1_solve()))
    if "--objects" in sys.argv:        # C4 / C5 described by objects: Kottke-smoothed on the device, no host eps array
        print(json.dumps(measure("C4 sphere 512^3 from objects (1 GPU)", workloads.c4_objects(), steps=20, kry=5)))
        print(json.dumps(measure("C5 metalens 1024x1024x96 from objects (1 GPU)", workloads.c5_objects(), steps=20, kry=5)))
    if "--c4" in sys.argv:
        t0 = time.perf_counter()
        w4 = workloads.c4_scatterer()
        sys.stderr.write(f"C4 eps built in {time.perf_counter() - t0:.1f} s\n")
        print(json.dumps(measure("C4 dielectric sphere 512^3 (1 GPU)", w4, steps=20, kry=5)))
        del w4
    if "--c5" in sys.argv:
        t0 = time.perf_counter()
        w5 = workloads.c5_metalens()
        sys.stderr.write(f"C5 eps built in {time.perf_counter() - t0:.1f} s\n")
        print(json.dumps(measure("C5 metalens 1024x1024x96 (weak-scaling unit, 1 GPU)", w5, steps=20, kry=5)))
