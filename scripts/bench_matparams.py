"""Throughput of the material pipeline kernel (fdfd_calc_matparams) on one GPU: cells/s and output
GB/s for (a) the C4 scatterer - one ball of radius 100 cells in a 512^3 box (z-slab of 128 planes), and (b) a C5-like
pillar field - 32x32 cylinders of seeded radii on a 1024x1024 grid (24 planes through the pillars).
Run on the GPU box:  python scripts/bench_matparams.py  (prints one JSON line per case; not part of bench.py)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maxwellfdm_jl_b200 as fb   # noqa: E402


def run(name, N, shapes, pinds, params, k0, k1, reps=3):
    lp = [np.arange(n + 1.0) for n in N]
    g = fb.Grid(lp, (True, True, False))
    best = 1e30
    for _ in range(reps):
        t = time.perf_counter()
        out = fb.calc_matparams_array(g, (fb.EE,) * 3, fb.EE, shapes, pinds, params, k0=k0, k1=k1, device=0,
                                      julia_layout=True)
        best = min(best, time.perf_counter() - t)
    cells = N[0] * N[1] * (k1 - k0)
    nint = int(np.count_nonzero(out[1, 0]))
    print(json.dumps({"case": name, "grid": list(N), "planes": [k0, k1], "cells": cells, "shapes": len(shapes),
                      "interface_corner_voxels": nint, "seconds_end_to_end": best, "mcells_s_end_to_end": cells / best / 1e6,
                      "note": "wall clock through the C ABI incl. the D2H copy of the 144 B/cell array; the kernel "
                              "itself is timed with ncu"}), flush=True)


def main():
    N = (512, 512, 512)
    shapes = [fb.Box([256, 256, 256], [512, 512, 512]), fb.Ball([256, 256, 256], 100.0)]
    run("C4 sphere", N, shapes, [0, 1], [np.eye(3), 4 * np.eye(3)], 192, 320)
    rng = np.random.default_rng(7)
    N = (1024, 1024, 96)
    shapes = [fb.Box([512, 512, 48], [1024, 1024, 96]), fb.Box([512, 512, 12], [1024, 1024, 12])]
    pinds = [0, 1]
    for j in range(32):
        for i in range(32):
            shapes.append(fb.Cylinder([16 + 32 * i, 16 + 32 * j, 44], 5 + 8 * rng.random(), 20, axis=2))
            pinds.append(2)
    run("C5 pillars", N, shapes, pinds, [np.eye(3), 2.1 * np.eye(3), 6 * np.eye(3)], 30, 54)


if __name__ == "__main__":
    main()
