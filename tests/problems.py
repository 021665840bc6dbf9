"""Seeded synthetic problems shared by the CPU and GPU tests (oracle side + product-side inputs)."""
import itertools

import numpy as np

from oracle.grid import Grid, EE, HH, create_stretched_dls, create_e_mikL
from oracle import operators as op
from oracle.matfree import MatFreeOperator

SEED = 20261017


def crandn(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


class Problem:
    """One operator instance: inputs in the form the C ABI takes + oracle builders."""

    def __init__(self, N, isbloch=(True, True, True), boundft=(EE, EE, EE), full_eps=False, with_mu=False,
                 ft=EE, omega=1.1 - 0.05j, npml=1, uniform=False, seed=SEED, cmpfirst=True, weighted_out=False,
                 kb_scale=1.0, full_mu=False):
        rng = np.random.default_rng(seed)
        self.N = tuple(int(n) for n in N)
        self.isbloch = tuple(bool(b) for b in isbloch)
        self.boundft = tuple(boundft)
        self.ft, self.omega, self.cmpfirst, self.weighted_out = ft, omega, cmpfirst, weighted_out
        if uniform:
            lprim = tuple(np.arange(n + 1, dtype=float) for n in self.N)
        else:
            lprim = tuple(np.concatenate(([0.0], np.cumsum(0.5 + rng.random(n)))) for n in self.N)
        self.grid = Grid(lprim, self.isbloch)
        Npml = (tuple(min(npml, n // 2) for n in self.N),) * 2
        self.sdl_e, self.sdl_m, self.sei, self.smi = create_stretched_dls(0.9 + 0.1j, self.grid, Npml, self.boundft)
        kb = np.where(self.isbloch, kb_scale * rng.random(3), 0.0)
        self.ph = create_e_mikL(kb, self.grid)
        self.eps = np.zeros(self.N + (3, 3), complex)
        self.mu = np.zeros(self.N + (3, 3), complex)
        for v in range(3):
            self.eps[..., v, v] = 2 + 0.3 * crandn(rng, *self.N)
            self.mu[..., v, v] = (1.5 + 0.2 * crandn(rng, *self.N)) if with_mu else 1.0
        if full_eps:
            for v, u in itertools.permutations(range(3), 2):
                self.eps[..., v, u] = 0.3 * crandn(rng, *self.N)
        if full_mu:   # only meaningful for ft == HH (mu is the mass parameter there, model.jl:238-240)
            for v, u in itertools.permutations(range(3), 2):
                self.mu[..., v, u] = 0.25 * crandn(rng, *self.N)
        self.with_mu, self.full_eps = with_mu or full_mu, full_eps
        self.n = 3 * int(np.prod(self.N))
        self.rng = rng

    # ---- oracle ---------------------------------------------------------------------------
    def oracle_csc(self):
        Ce, Cm = op.create_curls(self.sei, self.smi, self.boundft, self.isbloch, self.ph, self.cmpfirst)
        Pe, Pm = op.create_paramops(self.eps, self.mu, self.sdl_e, self.sdl_m, self.sei, self.smi, self.boundft,
                                    self.isbloch, self.ph, self.cmpfirst, self.weighted_out)
        return op.create_A(self.ft, self.omega, Pe, Pm, Ce, Cm), (Pe, Pm, Ce, Cm)

    def oracle_matfree(self):
        return MatFreeOperator(self.ft, self.omega, self.eps, self.mu if self.with_mu else None, self.sdl_e,
                               self.sdl_m, self.boundft, self.isbloch, self.ph, self.cmpfirst, self.weighted_out)

    def random_x(self, seed=1):
        return crandn(np.random.default_rng(SEED + seed), self.n)

    # ---- product --------------------------------------------------------------------------
    def operator(self, **kw):
        import maxwellfdm_jl_b200 as fb
        return fb.FdfdOperator(self.N, self.isbloch, self.sdl_e, self.sdl_m, self.omega, self.eps,
                               self.mu if self.with_mu else None, self.ph,
                               boundft=["E" if b == EE else "H" for b in self.boundft],
                               ft="E" if self.ft == EE else "H", order_cmpfirst=self.cmpfirst,
                               weighted_out_avg=self.weighted_out, **kw)


def matparams_scene(N, isbloch, uniform=False, nshape=6, aniso=False, seed=SEED):
    """Seeded random scene for the material pipeline (N4): a background box plus boxes / balls / cylinders with
    scalar or full-tensor materials; two consecutive objects share one material.  Returns (lprim, oracle shapes,
    product shapes, pinds, params)."""
    from oracle import matparams as omp
    import maxwellfdm_jl_b200 as fb
    rng = np.random.default_rng(seed)
    lp = [np.arange(n + 1.0) if uniform else np.concatenate(([0.0], np.cumsum(0.6 + 0.8 * rng.random(n)))) for n in N]
    Ls = [a[-1] for a in lp]
    o_sh, f_sh = [omp.Box([l / 2 for l in Ls], Ls)], [fb.Box([l / 2 for l in Ls], Ls)]
    params, pinds = [np.eye(3, dtype=complex)], [0]
    for s in range(nshape):
        c = [rng.random() * l for l in Ls]
        if s % 3 == 0:
            r = [0.5 + 2.5 * rng.random() for _ in range(3)]
            o_sh.append(omp.Box(c, r)); f_sh.append(fb.Box(c, r))
        elif s % 3 == 1:
            R = 1 + 2.5 * rng.random()
            o_sh.append(omp.Ball(c, R)); f_sh.append(fb.Ball(c, R))
        else:
            R, h, ax = 0.8 + 2 * rng.random(), 0.5 + 2 * rng.random(), int(rng.integers(3))
            o_sh.append(omp.Cylinder(c, R, h, ax)); f_sh.append(fb.Cylinder(c, R, h, ax))
        if s == 3:
            pinds.append(pinds[-1])
        else:
            P = np.eye(3) * (2 + 10 * rng.random()) + (0.3 * crandn(rng, 3, 3) if aniso else 0)
            params.append(P.astype(complex))
            pinds.append(len(params) - 1)
    return lp, o_sh, f_sh, pinds, params


MATPARAMS_CASES = [  # (N, isbloch, boundft, ft, uniform grid, number of shapes, anisotropic materials)
    ((12, 10, 9), (True, False, True), (EE, EE, EE), EE, True, 6, False),
    ((12, 10, 9), (False, True, False), (EE, HH, EE), EE, False, 6, True),
    ((13, 9, 11), (True, True, True), (EE, EE, EE), HH, False, 6, True),
    ((9, 17, 6), (False, False, False), (HH, HH, HH), EE, False, 9, False),
]


# ---- 2-D / 1-D models (ModelTE, ModelTM, ModelTEM) -----------------------------------------------------------------
REDUCED_CASES = [   # kind, N, isbloch, boundft, ft, order_cmpfirst
    ("TE", (9, 7), (True, False), (EE, EE), EE, True),
    ("TE", (12, 5), (False, True), (HH, EE), EE, False),
    ("TE", (8, 8), (True, True), (EE, HH), HH, True),
    ("TE", (70, 33), (False, False), (EE, EE), EE, True),
    ("TM", (9, 7), (True, False), (EE, EE), EE, True),
    ("TM", (6, 11), (False, False), (HH, HH), HH, False),
    ("TM", (35, 10), (True, True), (EE, HH), HH, True),
    ("TEM", (17,), (False,), (EE,), EE, True),
    ("TEM", (9,), (True,), (HH,), HH, True),
]


def reduced_model_check(fb, kind, N, isbloch, boundft, ft, cmpfirst, device=0, seed=5):
    """Reference call sequence on a ModelTE / ModelTM / ModelTEM (random non-uniform grid, PML on the non-periodic
    axes, Bloch phases, full 2x2 tensors where the formulation allows them, point source + random currents) against
    the K-dimensional oracle (oracle/reduced.py).  Returns the relative errors of A x, A^T x, b, the post-processed
    field and the solved field."""
    import scipy.sparse.linalg as spla
    from oracle import reduced as ored
    rng = np.random.default_rng(seed)
    kd = getattr(ored, kind)
    lprim = [np.concatenate(([0.0], np.cumsum(0.5 + rng.random(n)))) for n in N]
    mdl = {"TE": fb.ModelTE, "TM": fb.ModelTM, "TEM": fb.ModelTEM}[kind](fb.Grid(lprim, isbloch))
    fb.set_boundft(mdl, boundft)
    fb.set_wpml(mdl, 1.3)
    fb.set_Npml(mdl, ([1 if not b and n > 4 else 0 for b, n in zip(isbloch, N)],) * 2)
    fb.set_kbloch(mdl, [0.3 * b for b in isbloch])
    mdl.order_cmpfirst = cmpfirst

    def rand_param(Kf, diag):
        P = np.zeros(tuple(N) + (Kf, Kf), complex)
        for i in range(Kf):
            P[..., i, i] = 1.5 + rng.random(N) + 0.1j * rng.random(N)
        if not diag:
            for i, j in itertools.permutations(range(Kf), 2):
                P[..., i, j] = 0.2 * (rng.random(N) - 0.5) + 0.05j * rng.random(N)
        return P

    mdl.eps_arr[...] = rand_param(len(kd["cmp_e"]), diag=ft == HH)     # the divided tensor must be diagonal (model.jl:236,239)
    mdl.mu_arr[...] = rand_param(len(kd["cmp_m"]), diag=ft == EE)
    mdl.je_arr[...] = crandn(rng, *mdl.je_arr.shape)
    mdl.jm_arr[...] = crandn(rng, *mdl.jm_arr.shape)
    centre = [0.5 * (a[0] + a[-1]) + 0.1 for a in lprim]
    fb.add_srce(mdl, fb.PointSrc(centre, [1.0] * len(kd["cmp_e"])))
    w = 1.1 - 0.2j
    Ps, Cs, js = fb.create_paramops(mdl), fb.create_curls(mdl), fb.create_srcs(mdl)
    A, b = fb.create_linsys(ft, w, Ps, Cs, js, device=device)
    sdl_e, sdl_m, _, _ = fb.create_stretched_dls(mdl)
    S = ored.ReducedSystem(kd, mdl.eps_arr, mdl.mu_arr, sdl_e, sdl_m, boundft, isbloch, fb.create_e_mikL(mdl), cmpfirst)
    Ar = S.A(ft, w)
    assert A.n == Ar.shape[0] == mdl.length(ft)
    x = crandn(rng, A.n)
    errs = {"apply": rel(A @ x, Ar @ x), "transpose": rel(A.rmatvec_T(x), Ar.T @ x), "b": rel(b, S.b(ft, w, *js))}
    if ft == EE:
        errs["post"] = rel(fb.h_from_e(x, w, Ps, Cs, js), S.h_from_e(x, w, js[1]))
    else:
        errs["post"] = rel(fb.e_from_h(x, w, Ps, Cs, js), S.e_from_h(x, w, js[0]))
    xs, info = fb.solve(A, b, rtol=1e-11, maxit=5000)
    assert info["converged"], info
    errs["solve"] = rel(xs, spla.splu(Ar).solve(b))
    A.close()
    return errs
