"""TEST INFRASTRUCTURE ONLY - randomised scenes for the material pipeline kernel (CPU logic-check build) against the
oracle: shapes that stick out of the domain (ghost corners wrap / mirror), tiles at Bloch edges, more shapes than
the per-tile list holds (overflow path), tiny and odd grids, eps and mu locations, every boundft.

    FDFD_B200_LIB=build/emu/libfdfd_emu.so python tests/emu/fuzz_matparams_emu.py SEED NCASES
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

from oracle import matparams as omp                 # noqa: E402
from oracle.grid import Grid as OGrid               # noqa: E402
import maxwellfdm_jl_b200 as fb                     # noqa: E402


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def main():
    seed, ncases = int(sys.argv[1]), int(sys.argv[2])
    assert "EMULATED" in fb._lib.lib().fdfd_version().decode()
    rng = np.random.default_rng(seed)
    for case in range(ncases):
        N = tuple(int(v) for v in rng.choice([1, 2, 3, 5, 8, 9, 10, 17], 3))
        isbloch = tuple(bool(b) for b in rng.integers(0, 2, 3))
        boundft = tuple(int(b) for b in rng.integers(0, 2, 3))
        ft = int(rng.integers(0, 2))
        uniform = bool(rng.integers(0, 2))
        lp = [np.arange(n + 1.0) - 1.5 if uniform else np.concatenate(([0.0], np.cumsum(0.6 + 0.8 * rng.random(n)))) - 2.0 for n in N]
        lo = [a[0] for a in lp]
        Ls = [a[-1] - a[0] for a in lp]
        many = case % 7 == 3                                  # > 512 shapes: the per-tile list overflows
        nshape = 600 if many else int(rng.integers(1, 9))
        o_sh = [omp.Box([l + L / 2 for l, L in zip(lo, Ls)], [2 * L for L in Ls])]
        f_sh = [fb.Box([l + L / 2 for l, L in zip(lo, Ls)], [2 * L for L in Ls])]
        params, pinds = [np.eye(3, dtype=complex) * (1 + rng.random())], [0]
        aniso = bool(rng.integers(0, 2))
        for s in range(nshape):
            c = [l - 0.5 + rng.random() * (L + 1.0) for l, L in zip(lo, Ls)]       # centres may lie outside the domain
            kind = int(rng.integers(0, 3))
            scale = 0.3 if many else 1.0
            if kind == 0:
                r = [scale * (0.3 + 2.5 * rng.random()) for _ in range(3)]
                o_sh.append(omp.Box(c, r)); f_sh.append(fb.Box(c, r))
            elif kind == 1:
                R = scale * (0.4 + 2.5 * rng.random())
                o_sh.append(omp.Ball(c, R)); f_sh.append(fb.Ball(c, R))
            else:
                R, h, ax = scale * (0.4 + 2 * rng.random()), scale * (0.3 + 2 * rng.random()), int(rng.integers(3))
                o_sh.append(omp.Cylinder(c, R, h, ax)); f_sh.append(fb.Cylinder(c, R, h, ax))
            if rng.random() < 0.3 and len(params) > 1:
                pinds.append(int(rng.integers(1, len(params))))                     # re-use a material
            else:
                P = np.eye(3) * (1.5 + 10 * rng.random())
                if aniso:
                    Q = 0.3 * (rng.standard_normal((3, 3)) + 1j * rng.standard_normal((3, 3)))
                    P = P + (Q + Q.T if rng.integers(0, 2) else Q)
                params.append(P.astype(complex))
                pinds.append(len(params) - 1)
        tag = f"case {case}: N={N} bloch={isbloch} boundft={boundft} ft={ft} uniform={uniform} shapes={nshape} aniso={aniso}"
        ref = omp.calc_matparams(OGrid(lp, isbloch), boundft, ft, o_sh, pinds, params)
        got = fb.calc_matparams_array(fb.Grid(lp, isbloch), boundft, ft, f_sh, pinds, params, device=0)
        e = rel(got, ref)
        bad = np.abs(got - ref).max()
        if not (e < 1e-9):
            idx = np.unravel_index(np.argmax(np.abs(got - ref)), got.shape)
            print("FAIL", tag, "rel", e, "max", bad, "at", idx, got[idx], ref[idx], flush=True)
            sys.exit(1)
    print(f"matparams fuzz seed {seed}: {ncases} cases ok")


if __name__ == "__main__":
    main()
