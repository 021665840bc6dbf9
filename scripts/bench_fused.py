"""Apply throughput of the configurations that take the fused full-tensor kernel (dense off-diagonal blocks)."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import workloads  # noqa: E402

PEAK = 6549.8


def measure(name, w, steps=30):
    A = workloads.make_operator(w, device=0)
    n = A.n
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(n, 2, device="cuda", dtype=torch.float64, generator=g).view(torch.complex128).reshape(-1)
    y = torch.empty_like(x)
    A.bench_apply(x, y, warmup=5, iters=1)
    ms, _ = A.bench_apply(x, y, warmup=0, iters=steps)
    ms /= steps
    off = A.offdiag_fraction
    bpd = 48 + 32 * off
    print(json.dumps({"config": name, "gdof_s": round(n / ms / 1e6, 2), "f_off": round(off, 3),
                      "hbm_frac": round(bpd * n / (ms * 1e-3) / 1e9 / PEAK, 3), "launches_per_apply": A.launch_count / (steps + 6)}),
          flush=True)
    A.close()


if __name__ == "__main__":
    torch.cuda.set_device(0)
    w = workloads.c2_waveguide((200, 200, 200))
    rng = np.random.default_rng(7)
    for (v, u) in ((0, 1), (0, 2), (1, 2)):
        pert = 0.05 * (rng.random(w["eps"].shape[:3]) - 0.5)
        w["eps"][..., v, u] = pert
        w["eps"][..., u, v] = pert
    measure("C2 dense off-diagonal", w)
    measure("C5 metalens 512x512x96", workloads.c5_metalens((512, 512, 96)))
    measure("C3 PhC slab", workloads.c3_phc_slab())
