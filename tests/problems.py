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
                 kb_scale=1.0, full_mu=False, real_mass=False, sym_real_off=False):
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
        if real_mass:
            # lossless medium at a real frequency: the diagonal mass entries -w^2 eps_vv are real (the kernel then
            # streams them as doubles, apply_rowpair.cu MDR); PML stretch, Bloch phases and off-diagonals stay complex
            self.omega = omega = float(np.real(omega))
            mass = self.eps if ft == EE else self.mu
            for v in range(3):
                mass[..., v, v] = np.real(mass[..., v, v])
        if sym_real_off:
            # symmetric tensor with real off-diagonal entries (what subpixel smoothing of lossless reciprocal media gives);
            # together with real_mass the operator qualifies for the fused full-tensor shape of the row-pair kernel
            mass = self.eps if ft == EE else self.mu
            for v in range(3):
                for u in range(v + 1, 3):
                    mass[..., v, u] = np.real(mass[..., v, u])
                    mass[..., u, v] = mass[..., v, u]
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
    # create_Mcs (model.jl:287-306) where it is defined for a reduced model: the field whose components are the grid axes
    # (E of TE, H of TM) - component w averaged along its own axis w with the weights of model.jl:302-303
    Mce, Mcm = fb.create_Mcs(A)
    for cmps, Mc, isfwd, dl, dlo in ((kd["cmp_e"], Mce, [b_ != EE for b_ in boundft], sdl_m, sdl_e),
                                     (kd["cmp_m"], Mcm, [b_ != HH for b_ in boundft], sdl_e, sdl_m)):
        if tuple(cmps) != tuple(kd["cmp_s"]):
            continue
        ph = fb.create_e_mikL(mdl)
        f = crandn(rng, len(cmps) * int(np.prod(N)))
        F = op.field_vec2arr(f, N, len(cmps), cmpfirst) if len(N) == 3 else fb.field_vec2arr(f, N, cmpfirst)
        G = np.empty_like(F)
        for k in range(len(cmps)):
            M1 = op.create_m(k, bool(isfwd[k]), N, dl[k], 1 / np.asarray(dlo[k]), bool(isbloch[k]), ph[k]).to_scipy()
            G[..., k] = (M1 @ F[..., k].ravel(order="F")).reshape(N, order="F")
        errs["corners"] = rel(Mc(f), fb.field_arr2vec(G, cmpfirst))
    xs, info = fb.solve(A, b, rtol=1e-11, maxit=5000)
    assert info["converged"], info
    errs["solve"] = rel(xs, spla.splu(Ar).solve(b))
    A.close()
    return errs


def reduced_objects_check(fb, device=0):
    """add_obj / calc_matparams on ModelTE, ModelTM, ModelTEM.  (a) against the (numpy, 3-D) oracle on the scene the
    K-dimensional one stands for (shapes invariant along the missing axes, one periodic cell there); (b) analytic,
    independent of any 3-D code: across a planar x-normal interface on a uniform grid the tangential entries are the
    arithmetic mean and the normal entry the harmonic mean over the voxel's own x-extent.  Returns the number of checks."""
    from oracle import matparams as omp
    n = 0
    lx, ly = np.arange(13.0) - 2.0, np.arange(9.0) - 1.5
    e1, e2, x0 = 2.0, 9.0, 4.3          # eps = e2 for x < x0, e1 elsewhere
    for isbloch in ((False, False), (True, True)):
        res = {}
        for name, ctor in (("TE", fb.ModelTE), ("TM", fb.ModelTM)):
            mdl = ctor(fb.Grid([lx, ly], isbloch))
            fb.add_obj(mdl, "bg", fb.Box([4.0, 2.5], [50.0, 50.0]), eps=e1)
            fb.add_obj(mdl, "slab", fb.Box([x0 - 20.0, 2.5], [20.0, 50.0]), eps=e2)
            fb.create_paramops(mdl, device=device)             # runs calc_matparams (model.jl:143)
            res[name] = mdl
            assert np.array_equal(mdl.mu_arr[..., 0, 0], np.ones(mdl.grid.N))
        te, tm = res["TE"], res["TM"]
        xp = lx[:-1]                                             # primal planes; dual points sit half a cell further
        f_ex = np.clip((x0 - xp) / 1.0, 0, 1)                    # E_x: x-dual location, voxel [xp_i, xp_i + 1]
        f_ey = np.clip((x0 - (xp - 0.5)) / 1.0, 0, 1)            # E_y, E_z: x-primal location, voxel [xp_i - 1/2, xp_i + 1/2]
        # (the voxel around the first primal plane reaches across the boundary: its ghost half follows the pipeline's
        # boundary rule, which (a) below covers; the analytic statement is made for the voxels inside the domain)
        assert np.allclose(te.eps_arr[:, :, 0, 0], (1 / (f_ex / e2 + (1 - f_ex) / e1))[:, None], rtol=1e-12, atol=0)   # normal: harmonic
        assert np.allclose(te.eps_arr[1:, :, 1, 1], (f_ey * e2 + (1 - f_ey) * e1)[1:, None], rtol=1e-12, atol=0)        # tangential
        assert np.allclose(tm.eps_arr[:, :, 0, 0], te.eps_arr[:, :, 1, 1], rtol=1e-12, atol=0)     # E_z: tangential, same extent
        assert not te.eps_arr[:, :, 0, 1].any() and not te.eps_arr[:, :, 1, 0].any()
        n += 4
    # (a) disc + rotated-tensor rectangle on a non-uniform grid, every boundft, against the 3-D oracle
    rng = np.random.default_rng(9)
    lp = [np.concatenate(([0.0], np.cumsum(0.6 + 0.8 * rng.random(m)))) - 2.0 for m in (11, 9)]
    Q = np.array([[5.0, 0.7, 0.0], [0.7, 3.0, 0.0], [0.0, 0.0, 7.0]])
    for boundft in ((EE, EE), (HH, EE), (EE, HH)):
        for name, ctor, cmp_e, cmp_m in (("TE", fb.ModelTE, (0, 1), (2,)), ("TM", fb.ModelTM, (2,), (0, 1))):
            mdl = ctor(fb.Grid(lp, (True, False)))
            fb.set_boundft(mdl, boundft)
            fb.add_obj(mdl, "bg", fb.Box([2.0, 2.0], [40.0, 40.0]), eps=1.0, mu=1.0)
            fb.add_obj(mdl, "disc", fb.Ball([1.7, 2.2], 2.1), eps=Q, mu=2.0)
            fb.add_obj(mdl, "bar", fb.Box([4.1, 0.4], [1.3, 0.9]), eps=11.0, mu=[1.0, 3.0, 2.0])
            fb.calc_matparams(mdl, device=device)
            g3 = Grid([lp[0], lp[1], np.array([0.0, 1.0])], (True, False, True))
            big = 1e6 * max(1.0, max(mdl.grid.L))
            o_sh = [omp.Box([2.0, 2.0, 0.5], [40.0, 40.0, big]), omp.Cylinder([1.7, 2.2, 0.5], 2.1, big, 2),
                    omp.Box([4.1, 0.4, 0.5], [1.3, 0.9, big])]
            for arr, ft, cmps, prm in ((mdl.eps_arr, EE, cmp_e, [np.eye(3), Q, 11.0 * np.eye(3)]),
                                       (mdl.mu_arr, HH, cmp_m, [np.eye(3), 2.0 * np.eye(3), np.diag([1.0, 3.0, 2.0])])):
                prm3 = []
                for P in prm:
                    R = np.eye(3, dtype=complex)
                    R[np.ix_(cmps, cmps)] = np.asarray(P)[np.ix_(cmps, cmps)]
                    prm3.append(R)
                ref = omp.calc_matparams(g3, boundft + (EE,), ft, o_sh, [0, 1, 2], prm3, field_ortho_shape=len(cmps) == 1)
                ref = ref[:, :, 0][..., list(cmps), :][..., list(cmps)]
                assert rel(arr, ref) < 1e-10, (name, boundft, ft, rel(arr, ref))
                n += 1
    # 1-D: interval of eps 4 in vacuum, arithmetic averages (E_x, H_y orthogonal to z)
    lz = np.arange(11.0)
    tem = fb.ModelTEM(fb.Grid([lz], (False,)))
    fb.add_obj(tem, "bg", fb.Box([5.0], [50.0]), eps=1.0)
    fb.add_obj(tem, "film", fb.Box([5.15], [2.0]), eps=4.0)       # [3.15, 7.15]
    fb.calc_matparams(tem, device=device)
    zc = lz[:-1]                                                   # E_x on primal z: voxel [z - 1/2, z + 1/2]
    f = np.clip(np.minimum(zc + 0.5, 7.15) - np.maximum(zc - 0.5, 3.15), 0, 1)
    assert np.allclose(tem.eps_arr[:, 0, 0], 1.0 + 3.0 * f, rtol=1e-12, atol=0)
    n += 1
    with pytest_raises(ValueError):
        fb.add_obj(tem, "bad", fb.Box([1.0, 2.0], [1.0, 1.0]), eps=2.0)
        fb.calc_matparams(tem, device=device)
    return n + 1


class pytest_raises:
    """minimal stand-in for pytest.raises (this module is also imported by the stand-alone runners of tests/emu)"""

    def __init__(self, exc):
        self.exc = exc

    def __enter__(self):
        return self

    def __exit__(self, et, ev, tb):
        if et is None:
            raise AssertionError(f"{self.exc.__name__} not raised")
        return issubclass(et, self.exc)


REDUCED_PATTERN_CASES = REDUCED_CASES + [("TE", (1, 3), (True, True), (EE, EE), EE, True),
                                         ("TM", (2, 2), (True, False), (HH, EE), EE, True),
                                         ("TEM", (1,), (True,), (EE,), EE, True), ("TEM", (2,), (True,), (EE,), HH, False)]


def reduced_pattern_check(fb, kind, N, isbloch, boundft, ft, cmpfirst, device=-2, seed=3):
    """CSC index pattern of a 2-D / 1-D model's A (ReducedOperator.export_pattern: the block of the 3-D debug export)
    against the Julia-structure K-dimensional oracle (oracle/reduced.julia_csc): colptr / rowval bit-exact, values
    relative to the largest entry.  Also ties julia_csc to the scipy assembly of the same oracle."""
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
                P[..., i, j] = 0.2 * (rng.random(N) - 0.5)
        return P

    mdl.eps_arr[...] = rand_param(len(kd["cmp_e"]), ft == HH)
    mdl.mu_arr[...] = rand_param(len(kd["cmp_m"]), ft == EE)
    w = 1.1 - 0.2j
    A = fb.create_A(ft, w, mdl, device=device)
    colptr, rowval, nz = A.export_pattern()
    A.close()
    sdl_e, sdl_m, _, _ = fb.create_stretched_dls(mdl)
    args = (mdl.eps_arr, mdl.mu_arr, sdl_e, sdl_m, boundft, isbloch, fb.create_e_mikL(mdl), cmpfirst)
    J = ored.julia_csc(kd, ft, w, *args)
    cp, rv = J.julia_pattern()
    assert colptr.dtype == np.int64 and rowval.dtype == np.int64
    assert np.array_equal(colptr, cp) and np.array_equal(rowval, rv), (kind, N, isbloch, boundft, ft, cmpfirst)
    S = ored.ReducedSystem(kd, *args).A(ft, w)
    assert abs(J.to_scipy() - S).max() <= 1e-13 * abs(S).max()
    return float(np.abs(nz - J.nzval).max() / np.abs(J.nzval).max())
