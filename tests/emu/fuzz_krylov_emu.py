"""TEST INFRASTRUCTURE ONLY - BiCGSTAB / QMR trajectories of the library (CPU logic-check build) against textbook numpy
iterations on the oracle operator: after K iterations the iterates must agree (wrong inner products - e.g. in the
apply-epilogue fusion or the correction-pass deltas - keep r = b - A x consistent, so only a trajectory comparison
sees them).

    FDFD_B200_LIB=build/emu/libfdfd_emu.so python tests/emu/fuzz_krylov_emu.py SEED NCASES
"""
import ctypes as C
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


def bicgstab_ref(A, b, K):
    x = np.zeros_like(b)
    r = b.copy()
    rh = r.copy()
    p = r.copy()
    rho = np.vdot(rh, r)
    for _ in range(K):
        v = A(p)
        alpha = rho / np.vdot(rh, v)
        s = r - alpha * v
        t = A(s)
        omega = np.vdot(t, s) / np.vdot(t, t)
        x = x + alpha * p + omega * s
        r = s - omega * t
        rho_new = np.vdot(rh, r)
        beta = (rho_new / rho) * (alpha / omega)
        p = r + beta * (p - omega * v)
        rho = rho_new
    return x


def main():
    seed, ncases = int(sys.argv[1]), int(sys.argv[2])
    assert "EMULATED" in L.lib().fdfd_version().decode()
    rng = np.random.default_rng(seed)
    for case in range(ncases):
        N = (int(rng.choice([5, 9, 31, 33, 40])), int(rng.choice([4, 7, 13, 20])), int(rng.choice([3, 6, 11, 17])))
        isbloch = tuple(bool(b) for b in rng.integers(0, 2, 3))
        ft = int(rng.integers(0, 2))
        full = bool(rng.integers(0, 2))
        boundft = (EE,) * 3 if rng.integers(0, 2) else tuple(int(b) for b in rng.integers(0, 2, 3))
        kw = dict(full_eps=full and ft == EE, full_mu=full and ft == HH, with_mu=bool(rng.integers(0, 2)) or ft == HH)
        # a third of the cases: lossless medium at a real frequency (real material rows; with a full tensor half of them real
        # and symmetric: the fused shape of the row-pair kernel with the fused dots in its epilogue)
        if rng.integers(0, 3) == 0:
            kw.update(real_mass=True, sym_real_off=bool(full and rng.integers(0, 2)))
        p = Problem(N, isbloch, boundft, ft=ft, omega=1.2 - 0.3j, seed=int(rng.integers(1 << 30)), **kw)
        mass = p.eps if ft == EE else p.mu
        if full and rng.integers(0, 2):          # sparse off-diagonals -> diagonal kernel + correction pass (+ dot deltas)
            z0 = int(rng.integers(0, N[2]))
            for v, u in itertools.permutations(range(3), 2):
                mass[:, :, :z0, v, u] = 0
                mass[:, :, z0 + 1:, v, u] = 0
                mass[: N[0] // 2, :, :, v, u] = 0
        mf = p.oracle_matfree()
        b = p.random_x(3)
        K = 5
        A = p.operator(device=0, kernel=2)
        tag = (f"case {case}: N={N} bloch={isbloch} boundft={boundft} ft={ft} full={full} offfrac={A.offdiag_fraction:.2f} "
               f"real={kw.get('real_mass', False)} symreal={kw.get('sym_real_off', False)}")
        x = np.zeros(A.n, complex)
        iters, relres = C.c_int(), C.c_double()
        code = L.lib().fdfd_solve(A._h, L.BICGSTAB, b.ctypes.data, x.ctypes.data, L.DEVICE, 1e-300, K, 1, C.byref(iters),
                                  C.byref(relres), None)
        assert code in (L.OK, L.ENOCONV), (tag, code)
        xr = bicgstab_ref(mf, b, K)
        e = rel(x, xr)
        true_res = rel(mf(x), b)
        A.close()
        tol = 1e-9
        if e >= tol:
            # a lossless medium at a real frequency makes the iteration itself ill-conditioned (five steps can amplify the
            # last bit by 1e7): judge the trajectory by what the GENERAL kernel - different summation order, same
            # arithmetic - does on the same problem
            An = p.operator(device=0, kernel=1)
            xn = np.zeros(An.n, complex)
            code = L.lib().fdfd_solve(An._h, L.BICGSTAB, b.ctypes.data, xn.ctypes.data, L.DEVICE, 1e-300, K, 1, C.byref(iters),
                                      C.byref(iters_relres := C.c_double()), None)
            An.close()
            tol = max(tol, 20 * rel(xn, xr))
        if not (e < tol and abs(true_res - relres.value) < 1e-9 * max(1.0, true_res)):
            print("FAIL", tag, "trajectory", e, "tol", tol, "relres", relres.value, "true", true_res, flush=True)
            sys.exit(1)
    print(f"krylov fuzz seed {seed}: {ncases} cases ok")


if __name__ == "__main__":
    main()
