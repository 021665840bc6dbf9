"""N3 row measurement: the HH formulation (model.jl:238-240) and the dual arrangement (boundft all-HH) on the C2 grid.
Prints one JSON line per variant (apply GDOF/s with CUDA events inside fdfd_bench_apply; BiCGSTAB it/s)."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import workloads  # noqa: E402
import maxwellfdm_jl_b200 as fb  # noqa: E402

PEAK = 6549.8


def main():
    torch.cuda.set_device(0)
    w = workloads.c2_waveguide((200, 200, 200))
    eps_full = w["eps"]
    eps_diag = eps_full.copy()
    for v in range(3):
        for u in range(3):
            if u != v:
                eps_diag[..., v, u] = 0
    n = 3 * 200 ** 3
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(n, 2, device="cuda", dtype=torch.float64, generator=g).view(torch.complex128).reshape(-1)
    y = torch.empty_like(x)
    variants = (
        ("EE default (forward kernel), diagonal eps", dict(ft="E"), eps_diag, False, 0),
        ("HH formulation (mirrored kernel), diagonal eps", dict(ft="H"), eps_diag, False, 0),
        ("HH formulation, general kernel", dict(ft="H"), eps_diag, False, 1),
        ("EE on boundft all-HH (mirrored kernel), full eps", dict(ft="E", boundft=["H"] * 3), eps_full, True, 0),
        ("EE on boundft all-HH, general kernel", dict(ft="E", boundft=["H"] * 3), eps_full, True, 1),
    )
    only = int(sys.argv[sys.argv.index("--variant") + 1]) if "--variant" in sys.argv else None
    for iv, (name, kw, eps, off, kernel) in enumerate(variants):
        if only is not None and iv != only:
            continue
        A = fb.FdfdOperator(w["N"], w["isbloch"], w["sdl_e"], w["sdl_m"], w["omega"], eps, None, w["e_mikL"],
                            device=0, kernel=kernel, eps_has_offdiag=off, **kw)
        tot, mn = A.bench_apply(x, y, warmup=5, iters=50)
        t = tot / 50 * 1e-3            # fdfd_bench_apply reports milliseconds
        f = A.offdiag_fraction if off else 0.0
        bpd = 48 + 32 * f
        line = {"variant": name, "gdof_s": n / t / 1e9, "us": t * 1e6, "bytes_per_dof": bpd,
                "hbm_frac": n * bpd / t / 1e9 / PEAK}
        if kernel == 0 and only is None:
            b = torch.randn(n, 2, device="cuda", dtype=torch.float64, generator=g).view(torch.complex128).reshape(-1)
            xs = torch.zeros_like(b)
            ts = A.bench_solve(b, xs, "bicgstab", warmup=2, iters=40)
            line["bicgstab_it_s"] = 40 / (ts * 1e-3)
        print(json.dumps(line), flush=True)
        A.close()


if __name__ == "__main__":
    main()
