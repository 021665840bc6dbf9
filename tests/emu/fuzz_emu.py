"""TEST INFRASTRUCTURE ONLY - randomised campaign of the tiled kernel under the CPU logic-check build against the
matrix-free oracle: random grid sizes (incl. sizes just around the tile and chunk boundaries), boundary conditions,
arrangements, layouts, material kinds, sparse / dense off-diagonal patterns, forced z-chunk lengths (FDFD_LZ) and,
per process, a forced tile height (FDFD_TY).  Since round 2 half of the cases have real diagonal mass entries (MDR shape of
the row-pair kernel) and half of those a real symmetric off-diagonal tensor (fused shape, or - below FDFD_RP_FUSE_MIN - the
two-pass plan on the row-pair kernel); per process the row-pair plan can be bent with FDFD_RP_NCHUNK / FDFD_RP_GRID /
FDFD_RP_FUSE_MIN / FDFD_RP_DEBUG=16|32|64.

    FDFD_B200_LIB=build/emu/libfdfd_emu.so [FDFD_TY=16] python tests/emu/fuzz_emu.py SEED NCASES
"""
import itertools
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

from oracle.grid import EE, HH                      # noqa: E402
from problems import Problem, rel                   # noqa: E402
import maxwellfdm_jl_b200 as fb                     # noqa: E402

L = fb._lib


def apply_dev(A, x, transpose=False):
    y = np.full(A.n, np.nan + 1j * np.nan)
    f = L.lib().fdfd_apply_transpose if transpose else L.lib().fdfd_apply
    L.check(f(A._h, x.ctypes.data, y.ctypes.data, L.DEVICE), A._h)
    return y


def main():
    seed, ncases = int(sys.argv[1]), int(sys.argv[2])
    assert "EMULATED" in L.lib().fdfd_version().decode()
    rng = np.random.default_rng(seed)
    nx_pool = [1, 2, 3, 7, 29, 30, 31, 32, 33, 59, 60, 61, 62, 64, 89, 91]
    ny_pool = [1, 2, 3, 5, 6, 7, 8, 11, 12, 13, 14, 15, 20, 27, 28, 29, 30]
    nz_pool = [1, 2, 3, 4, 5, 7, 9, 16, 23, 31, 32, 33, 41, 44, 63]
    for case in range(ncases):
        N = (int(rng.choice(nx_pool)), int(rng.choice(ny_pool)), int(rng.choice(nz_pool)))
        isbloch = tuple(bool(b) for b in rng.integers(0, 2, 3))
        boundft = tuple(int(b) for b in rng.integers(0, 2, 3))
        ft = int(rng.integers(0, 2))
        cmpfirst = bool(rng.integers(0, 2))
        full = bool(rng.integers(0, 2))
        with_mu = bool(rng.integers(0, 2))
        full_mass = full
        real_mass = bool(rng.integers(0, 2))
        sym_real = real_mass and full_mass and bool(rng.integers(0, 2))
        kw = dict(full_eps=full_mass and ft == EE, full_mu=full_mass and ft == HH, with_mu=with_mu or ft == HH,
                  real_mass=real_mass, sym_real_off=sym_real)
        p = Problem(N, isbloch, boundft, ft=ft, cmpfirst=cmpfirst, seed=int(rng.integers(1 << 30)), npml=int(rng.integers(0, 4)), **kw)
        mass = p.eps if ft == EE else p.mu
        pattern = int(rng.integers(0, 3))
        if full_mass and pattern > 0:      # sparse off-diagonals: a z range and an x range only, or a single plane
            z0, z1 = sorted(rng.integers(0, N[2] + 1, 2))
            if pattern == 2:
                z1 = min(z0 + 1, N[2])
            for v, u in itertools.permutations(range(3), 2):
                mass[:, :, :z0, v, u] = 0
                mass[:, :, z1:, v, u] = 0
                mass[: N[0] // 2, :, :, v, u] = 0
        lz = int(rng.choice([0, 1, 2, 3, 4, 5, 7, 13]))
        if lz:
            os.environ["FDFD_LZ"] = str(lz)
        else:
            os.environ.pop("FDFD_LZ", None)
        tag = (f"case {case}: N={N} bloch={isbloch} boundft={boundft} ft={ft} cmpfirst={cmpfirst} full={full_mass} mu={with_mu} "
               f"pattern={pattern} lz={lz} real_mass={real_mass} sym_real_off={sym_real}")
        try:
            A = p.operator(device=0, kernel=2)
            x = p.random_x()
            mf = p.oracle_matfree()
            e1 = rel(apply_dev(A, x), mf(x))
            An = p.operator(device=0, kernel=1)
            e2 = rel(apply_dev(A, x, True), apply_dev(An, x, True))
            A.close()
            An.close()
        except Exception as exc:  # noqa: BLE001
            print("FAIL", tag, repr(exc), flush=True)
            raise
        if not (e1 < 1e-12 and e2 < 1e-12):
            print("FAIL", tag, e1, e2, flush=True)
            sys.exit(1)
    print(f"fuzz seed {seed}: {ncases} cases ok (FDFD_TY={os.environ.get('FDFD_TY', '-')}, "
          f"{os.environ.get('FDFD_EMU_ASYNC', 'eager')}, shuffle={os.environ.get('FDFD_EMU_SHUFFLE', '0')})")


if __name__ == "__main__":
    main()
