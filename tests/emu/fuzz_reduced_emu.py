"""TEST INFRASTRUCTURE ONLY - randomised 2-D / 1-D models (ModelTE, ModelTM, ModelTEM on the 3-D kernels,
maxwellfdm.jl_b200/reduced.py) against the K-dimensional oracle: sizes around the tile edges, every boundft, Bloch /
symmetry mixes, both formulations and DOF orders; operator, right-hand side, post-processing, solve and the exported
index pattern.

    FDFD_B200_LIB=build/emu/libfdfd_emu.so python tests/emu/fuzz_reduced_emu.py SEED NCASES
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

from problems import reduced_model_check, reduced_pattern_check   # noqa: E402
import maxwellfdm_jl_b200 as fb                     # noqa: E402


def main():
    seed, ncases = int(sys.argv[1]), int(sys.argv[2])
    assert "EMULATED" in fb._lib.lib().fdfd_version().decode()
    rng = np.random.default_rng(seed)
    for case in range(ncases):
        kind = str(rng.choice(["TE", "TM", "TEM"]))
        K = 1 if kind == "TEM" else 2
        N = tuple(int(v) for v in rng.choice([1, 2, 3, 5, 8, 29, 30, 31, 33, 61], K))
        isbloch = tuple(bool(b) or n == 1 for b, n in zip(rng.integers(0, 2, K), N))   # a point source needs N > 1 on symmetry axes (source.jl:214)
        boundft = tuple(int(b) for b in rng.integers(0, 2, K))
        args = (kind, N, isbloch, boundft, int(rng.integers(0, 2)), bool(rng.integers(0, 2)))
        errs = reduced_model_check(fb, *args, seed=int(rng.integers(1 << 30)))
        if not (max(v for k, v in errs.items() if k != "solve") < 1e-12 and errs["solve"] < 1e-6):
            print("FAIL", args, errs, flush=True)
            sys.exit(1)
        ev = reduced_pattern_check(fb, *args, seed=int(rng.integers(1 << 30)))     # index pattern bit-exact (asserts inside)
        if not ev < 1e-13:
            print("FAIL pattern values", args, ev, flush=True)
            sys.exit(1)
    print(f"reduced-model fuzz seed {seed}: {ncases} cases ok")


if __name__ == "__main__":
    main()
