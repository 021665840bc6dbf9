#!/usr/bin/env python
"""ONE process, N GPUs: the single-call boundary (fdfd_multi_*, csrc/multi.cpp) against the oracle.

    python scripts/multi_check.py [NGPU ...]          default: 1 and every power of two up to the box's GPU count

For every NGPU: full-grid host vectors through MultiGpuOperator - apply and transposed apply on both DOF layouts, with and
without Bloch wrap in z, slabs shallow (plain staged path) and deep (sub-slab pipeline), full / diagonal / real material;
BiCGSTAB and QMR against a sparse direct solve; the model-level call create_A(ft, w, mdl, ngpu=N).  Prints
`MULTI_CHECK OK ...` and exits 0 when everything agrees (apply 1e-12, solve 1e-8)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import numpy as np
import scipy.sparse.linalg as spla
import torch

from problems import Problem, rel
import maxwellfdm_jl_b200 as fb


def multi_operator(p, ngpu, **kw):
    return fb.MultiGpuOperator(p.N, p.isbloch, p.sdl_e, p.sdl_m, p.omega, p.eps, p.mu if p.with_mu else None, p.ph,
                               boundft=["E" if b == 0 else "H" for b in p.boundft], ft="E" if p.ft == 0 else "H",
                               order_cmpfirst=p.cmpfirst, ngpu=ngpu, **kw)


def check(ngpu):
    fails, n = [], 0
    cases = [dict(N=(21, 18, 4 * ngpu + 3), isbloch=(True, True, True), full_eps=True, with_mu=True),
             dict(N=(21, 18, 4 * ngpu + 3), isbloch=(False, True, False), full_eps=True, cmpfirst=False),
             dict(N=(33, 10, 3 * ngpu + 1), isbloch=(True, False, True), boundft=(1, 1, 1), full_eps=True, with_mu=True),
             dict(N=(21, 18, 4 * ngpu + 1), isbloch=(True, True, False), ft=1, full_mu=True),
             dict(N=(40, 31, 17 * ngpu), isbloch=(True, False, True), full_eps=True),                       # pipelined
             dict(N=(40, 31, 16 * ngpu + 1), isbloch=(False, True, False), full_eps=True, real_mass=True, sym_real_off=True),
             dict(N=(40, 31, 17 * ngpu), isbloch=(False, False, True), full_eps=False, cmpfirst=False)]
    for cs in cases:
        p = Problem(**cs)
        A_ref, _ = p.oracle_csc()
        A = multi_operator(p, ngpu)
        x = p.random_x()
        e1 = rel(A @ x, A_ref.matvec(x))
        y = np.empty_like(x)
        e2 = rel(A.mul(y, x, transpose=True), A_ref.to_scipy().T @ x)
        A.close()
        n += 2
        if not (e1 < 1e-12 and e2 < 1e-12):
            fails.append((ngpu, cs, e1, e2))
    for ft, kw in ((0, dict(full_eps=True)), (1, dict(with_mu=True))):
        p = Problem((12, 10, 4 * ngpu + 1), (True, False, True), ft=ft, omega=1.3 - 0.4j, **kw)
        A_ref, _ = p.oracle_csc()
        b = A_ref.matvec(p.random_x(5))
        x_ref = spla.splu(A_ref.to_scipy().tocsc()).solve(b)
        A = multi_operator(p, ngpu)
        for method in ("bicgstab", "qmr"):
            xs, info = A.solve(b, method=method, rtol=1e-10, maxit=4000)
            e = rel(xs, x_ref)
            n += 1
            if not (e < 1e-7 and info["converged"]):
                fails.append((ngpu, "solve", ft, method, e, info))
        A.close()
    # the reference-level call sequence with ngpu=N (model.jl:209-246)
    from oracle.grid import EE
    N = (12, 9, 4 * ngpu + 2)
    g = fb.Grid(tuple(np.arange(n_ + 1.0) for n_ in N), (True, True, False))
    mdl = fb.ModelFull(g)
    fb.set_wpml(mdl, 0.9)
    fb.set_Npml(mdl, ((0, 0, 2), (0, 0, 2)))
    rng = np.random.default_rng(3)
    for v in range(3):
        mdl.eps_arr[..., v, v] = 2 + rng.random(N)
    fb.add_srce(mdl, fb.PointSrc([6.2, 4.1, N[2] / 2.0], [0, 0, 1], 1.0))
    A, b = fb.create_linsys(EE, 0.9, mdl, ngpu=ngpu)
    A1, b1 = fb.create_linsys(EE, 0.9, mdl)
    xv = np.random.default_rng(5).standard_normal(A.n) + 0j
    e = max(rel(b, b1), rel(A @ xv, A1 @ xv))
    n += 1
    if not e < 1e-13:
        fails.append((ngpu, "model-level create_linsys", e))
    A.close()
    A1.close()
    # objects instead of arrays (device_materials): every slab rasterises and smooths its own planes on its device
    mdl2 = fb.ModelFull(g)
    fb.set_wpml(mdl2, 0.9)
    fb.set_Npml(mdl2, ((0, 0, 2), (0, 0, 2)))
    fb.add_obj(mdl2, "bg", fb.Box([6.0, 4.5, N[2] / 2.0], [12.0, 9.0, float(N[2])]), eps=2.1)
    fb.add_obj(mdl2, "ball", fb.Ball([5.7, 4.2, N[2] / 2.0 + 0.3], 2.6), eps=11.7)
    Ps, Cs = fb.create_paramops(mdl2, device_materials=True), fb.create_curls(mdl2)
    Am = fb.create_A(EE, 0.9, Ps, Cs, ngpu=ngpu)
    As = fb.create_A(EE, 0.9, Ps, Cs)
    e = rel(Am @ xv, As @ xv)
    n += 1
    if not e < 1e-13:
        fails.append((ngpu, "objects through the single-call handle", e))
    Am.close()
    As.close()
    return n, fails


if __name__ == "__main__":
    ndev = torch.cuda.device_count()
    want = [int(a) for a in sys.argv[1:]] or [g for g in (1, 2, 4, 8) if g <= ndev]
    total, fails = 0, []
    for g in want:
        n, f = check(g)
        total += n
        fails += f
        print(f"ngpu {g}: {n} checks, {len(f)} failed", flush=True)
    for f in fails:
        print("FAIL", f)
    print(f"MULTI_CHECK {'OK' if not fails else 'FAILED'} ngpu {want} checks {total}")
    sys.exit(1 if fails else 0)
