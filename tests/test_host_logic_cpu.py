"""Host-side logic of the product (grid geometry, PML stretch, sources, DOF ordering, model API plumbing)
against the oracle and the reference's own source fixtures (test/source.jl) - no GPU needed."""
import numpy as np
import pytest

import maxwellfdm_jl_b200 as fb
from oracle import grid as og, source as osrc, operators as oop


def _rand_grid(rng, N, isbloch):
    lprim = tuple(np.concatenate(([0.0], np.cumsum(0.5 + rng.random(n)))) - 3.0 for n in N)
    return fb.Grid(lprim, isbloch), og.Grid(lprim, isbloch)


def test_grid_and_pml_match_oracle():
    rng = np.random.default_rng(5)
    for isbloch in ((True, False, True), (False, True, False)):
        g, o = _rand_grid(rng, (9, 7, 8), isbloch)
        assert g.N == o.N and np.allclose(g.L, o.L) and g.bounds == o.bounds
        for t in (0, 1):
            for w in range(3):
                assert np.array_equal(g.l[t][w], o.l[t][w]) and np.array_equal(g.dl[t][w], o.dl[t][w])
        for boundft in ((fb.EE,) * 3, (fb.HH, fb.EE, fb.HH)):
            mdl = fb.ModelFull(g)
            fb.set_wpml(mdl, 0.8 - 0.1j)
            fb.set_Npml(mdl, ((2, 0, 3), (1, 2, 0)))
            fb.set_boundft(mdl, boundft)
            fb.set_kbloch(mdl, (0.1, 0.0, -0.2))
            mine = fb.create_stretched_dls(mdl)
            ref = og.create_stretched_dls(0.8 - 0.1j, o, ((2, 0, 3), (1, 2, 0)), boundft)
            for a, b in zip(mine, ref):
                for w in range(3):
                    assert np.allclose(a[w], b[w], rtol=1e-15, atol=0)
            assert np.allclose(fb.create_e_mikL(mdl), og.create_e_mikL((0.1, 0.0, -0.2), o), rtol=1e-15)


def test_distweights_reference_table():
    """reference test/source.jl:17-60,64-104 (spot rows; the full table is in test_oracle_source.py)."""
    lp = np.cumsum(np.concatenate(([-1], np.arange(2, 23, 2))))
    lprim_g, ldual_g = 0.5 * (lp[:-1] + lp[1:]), lp[:-1]
    dom = (lprim_g[0], lprim_g[-1])
    lprim, dlprim, ldual, dldual = lprim_g[:-1], np.diff(ldual_g), ldual_g[1:], np.diff(lprim_g)
    r = (105 - 99) / (120 - 99)
    ind, wt = fb.distweights(105, fb.PRIM, dom, lprim, dlprim, True)
    assert ind == (9, 0) and np.allclose(wt, [1 / dlprim[9] * (1 - r), 1 / dlprim[0] * r])
    ind, wt = fb.distweights(0, fb.PRIM, dom, lprim, dlprim, False)
    assert ind == (1, 1) and wt == (0.0, 0.0)
    r = (1 - 0.5) / ((1 - 0) + (120 - 109))
    ind, wt = fb.distweights(0.5, fb.DUAL, dom, ldual, dldual, True)
    assert ind == (0, 9) and np.allclose(wt, [1 / dldual[0] * (1 - r), 1 / dldual[9] * r])
    with pytest.raises(ValueError):
        fb.distweights(-0.5, fb.DUAL, dom, ldual, dldual, False)
    with pytest.raises(ValueError):
        fb.distweights(1.0, fb.PRIM, (0, 2), [0], [9], False)


def test_distweights_and_sources_match_oracle_randomised():
    rng = np.random.default_rng(11)
    for isbloch in ((True, True, True), (False, False, False), (True, False, True)):
        g, o = _rand_grid(rng, (7, 6, 8), isbloch)
        for gt in (fb.PRIM, fb.DUAL):
            for w in range(3):
                for c in np.concatenate((rng.uniform(g.bounds[0][w], g.bounds[1][w], 20), g.l[gt][w][:3],
                                         [g.bounds[0][w], g.bounds[1][w]])):
                    a = fb.distweights(c, gt, (g.bounds[0][w], g.bounds[1][w]), g.l[gt][w], g.dl[gt][w], isbloch[w])
                    b = osrc.distweights(c, gt, (o.bounds[0][w], o.bounds[1][w]), o.l[gt][w], o.dl[gt][w], isbloch[w])
                    assert a[0] == b[0] and np.allclose(a[1], b[1], rtol=1e-14, atol=0)
        mdl = fb.ModelFull(g)
        c = [rng.uniform(g.bounds[0][w], g.bounds[1][w]) for w in range(3)]
        fb.add_srce(mdl, fb.PointSrc(c, [1, 2, -1], 0.3 + 0.1j))
        fb.add_srce(mdl, fb.PlaneSrc([0, 1, 0], c[1], [1, 0, 1], 2.0))
        fb.add_srcm(mdl, fb.PlaneSrc([0, 0, 1], c[2], [0, 1, 0]))
        je, jm = osrc.create_field_array(o.N), osrc.create_field_array(o.N)
        osrc.add_src(je, og.EE, (og.EE,) * 3, o, osrc.PointSrc(c, [1, 2, -1], 0.3 + 0.1j))
        osrc.add_src(je, og.EE, (og.EE,) * 3, o, osrc.PlaneSrc([0, 1, 0], c[1], [1, 0, 1], 2.0))
        osrc.add_src(jm, og.HH, (og.EE,) * 3, o, osrc.PlaneSrc([0, 0, 1], c[2], [0, 1, 0]))
        assert np.allclose(mdl.je_arr, je, rtol=1e-14, atol=0) and np.allclose(mdl.jm_arr, jm, rtol=1e-14, atol=0)
        vje, vjm = fb.create_srcs(mdl)
        assert np.array_equal(vje, oop.field_arr2vec(mdl.je_arr)) and vje is not mdl.je_arr
        fb.clear_srcs(mdl)
        assert not mdl.je_arr.any() and not mdl.jm_arr.any()


def test_pointsrc_conservation_exact():
    """reference test/source.jl:172-204: 8 non-zeros per component and exact sum on a uniform grid."""
    g = fb.Grid((np.arange(-10, 11.0),) * 3, (True, True, True))
    mdl = fb.ModelFull(g)
    src = fb.PointSrc([0.7, 0.7, 0.7], [1, 1, 1])
    fb.add_srce(mdl, src)
    for c in range(3):
        assert np.count_nonzero(mdl.je_arr[..., c]) == 8
        assert mdl.je_arr[..., c].sum() == src.Idr * src.p[c]
    with pytest.raises(ValueError):
        fb.PlaneSrc([1, 1, 0], 0, [1, 0, 0])


def test_dof_ordering_rule():
    """model.jl:75-83: r = c + 3*(i + Nx*(j + Ny*k)) (cmp-first) or i + Nx*(j + Ny*(k + Nz*c))."""
    N = (4, 3, 5)
    F = np.arange(np.prod(N) * 3).reshape(N + (3,)).astype(complex)
    v = fb.field_arr2vec(F, True)
    w = fb.field_arr2vec(F, False)
    for (i, j, k, c) in ((0, 0, 0, 0), (3, 2, 4, 2), (1, 2, 3, 1)):
        assert v[c + 3 * (i + N[0] * (j + N[1] * k))] == F[i, j, k, c]
        assert w[i + N[0] * (j + N[1] * (k + N[2] * c))] == F[i, j, k, c]
    assert np.array_equal(fb.field_vec2arr(v, N, True), F) and np.array_equal(fb.field_vec2arr(w, N, False), F)
    mdl = fb.ModelFull(fb.Grid(tuple(np.arange(n + 1.0) for n in N), (True,) * 3))
    assert mdl.size(fb.EE) == (3,) + N and mdl.length(fb.EE) == 3 * 60
    with pytest.raises(ValueError):
        fb.create_A(7, 1.0, mdl)      # reference: @error "ft = ... is unsupported." (model.jl:242)
    with pytest.raises(ValueError):
        fb.create_A(7, 1.0, fb.create_paramops(mdl), fb.create_curls(mdl))


def test_reference_call_sequence_descriptors():
    """create_paramops / create_curls (model.jl:141-175) return descriptions the reference-shaped create_A consumes;
    on a host-only handle (device = -2) the assembled pattern equals the oracle's create_A for both formulations."""
    from oracle.grid import Grid as OGrid, create_stretched_dls as o_sdls, EE as OEE, HH as OHH
    from oracle import operators as op
    N = (5, 4, 6)
    lp = tuple(np.arange(n + 1.0) for n in N)
    mdl = fb.ModelFull(fb.Grid(lp, (True, False, True)))
    fb.set_wpml(mdl, 0.9)
    fb.set_Npml(mdl, ((0, 1, 0), (0, 2, 0)))
    fb.set_kbloch(mdl, (0.3, 0.0, 0.2))
    rng = np.random.default_rng(5)
    for v in range(3):
        mdl.eps_arr[..., v, v] = 2 + rng.random(N)
        mdl.mu_arr[..., v, v] = 1 + rng.random(N)
    Ps, Cs = fb.create_paramops(mdl), fb.create_curls(mdl)
    assert isinstance(Ps[0], fb.ParamOp) and Ps[0].kind == "eps" and Ps[1].kind == "mu"
    assert isinstance(Cs[0], fb.CurlOp) and (Cs[0].kind, Cs[1].kind) == ("Ce", "Cm")
    og = OGrid(lp, (True, False, True))
    sdl_e, sdl_m, sei, smi = o_sdls(0.9, og, ((0, 1, 0), (0, 2, 0)))
    ph = fb.create_e_mikL(mdl)
    Ce, Cm = op.create_curls(sei, smi, (OEE,) * 3, og.isbloch, ph)
    Pe, Pm = op.create_paramops(mdl.eps_arr, mdl.mu_arr, sdl_e, sdl_m, sei, smi, (OEE,) * 3, og.isbloch, ph)
    for ft, oft in ((fb.EE, OEE), (fb.HH, OHH)):
        A = fb.create_A(ft, 1.3, Ps, Cs, device=-2)
        assert fb.create_A(ft, 1.3, Ps, Cs, device=-2) is A
        cp, rv, nz = A.export_pattern()
        ref = op.create_A(oft, 1.3, Pe, Pm, Ce, Cm)
        assert np.array_equal(cp, ref.julia_pattern()[0]) and np.array_equal(rv, ref.julia_pattern()[1])
        assert np.abs(nz - ref.nzval).max() <= 1e-13 * np.abs(ref.nzval).max()
        A.close()
    # descriptions taken from different model settings do not mix
    fb.set_kbloch(mdl, (0.1, 0.0, 0.2))
    with pytest.raises(ValueError):
        fb.create_A(fb.EE, 1.3, Ps, fb.create_curls(mdl), device=-2)


def test_create_A_follows_new_paramops_on_shared_curls():
    """The reference keeps create_paramops and create_curls apart so that Cs can be reused when only the material
    changes (model.jl:141-175).  Ps from a later create_paramops call must give an operator with the NEW material,
    never the operator cached for the earlier Ps (the cache was keyed on geometry only)."""
    from oracle.grid import Grid as OGrid, create_stretched_dls as o_sdls, EE as OEE
    from oracle import operators as op
    N = (4, 5, 3)
    lp = tuple(np.arange(n + 1.0) for n in N)
    mdl = fb.ModelFull(fb.Grid(lp, (True, True, True)))
    for v in range(3):
        mdl.eps_arr[..., v, v] = 2.0
        mdl.mu_arr[..., v, v] = 1.0
    Cs = fb.create_curls(mdl)
    Ps1 = fb.create_paramops(mdl)
    A1 = fb.create_A(fb.EE, 1.1, Ps1, Cs, device=-2)
    assert fb.create_A(fb.EE, 1.1, Ps1, Cs, device=-2) is A1
    nz1 = A1.export_pattern()[2]
    for v in range(3):
        mdl.eps_arr[..., v, v] = 5.0                      # the material changes, the curls do not
    Ps2 = fb.create_paramops(mdl)
    A2 = fb.create_A(fb.EE, 1.1, Ps2, Cs, device=-2)
    assert A2 is not A1
    nz2 = A2.export_pattern()[2]
    og = OGrid(lp, (True, True, True))
    sdl_e, sdl_m, sei, smi = o_sdls(0.0, og, ((0, 0, 0), (0, 0, 0)))
    ph = fb.create_e_mikL(mdl)
    Ce, Cm = op.create_curls(sei, smi, (OEE,) * 3, og.isbloch, ph)
    Pe, Pm = op.create_paramops(mdl.eps_arr, mdl.mu_arr, sdl_e, sdl_m, sei, smi, (OEE,) * 3, og.isbloch, ph)
    ref = op.create_A(OEE, 1.1, Pe, Pm, Ce, Cm)
    assert np.abs(nz2 - ref.nzval).max() <= 1e-13 * np.abs(ref.nzval).max()
    assert np.abs(nz1 - nz2).max() > 1.0                  # the first operator really held the old material
    assert fb.create_A(fb.EE, 1.1, Ps2, Cs, device=-2) is A2
    A1.close(); A2.close()


def test_reduced_models_layout_and_sources():
    """ModelTE / ModelTM / ModelTEM (te.jl:4-14, tm.jl:4-14, tem.jl:4-13): array shapes, DOF order with Kf components
    (model.jl:75-83), sources on K-dimensional grids (isfield˔shp default: orthogonal complement -> true) against the
    oracle's K-generic add_src."""
    rng = np.random.default_rng(3)
    g2, o2 = _rand_grid(rng, (7, 5), (True, False))
    te, tm = fb.ModelTE(g2), fb.ModelTM(g2)
    assert te.eps_arr.shape == (7, 5, 2, 2) and te.mu_arr.shape == (7, 5, 1, 1) and te.je_arr.shape == (7, 5, 2)
    assert tm.eps_arr.shape == (7, 5, 1, 1) and tm.mu_arr.shape == (7, 5, 2, 2) and tm.jm_arr.shape == (7, 5, 2)
    assert te.size(fb.EE) == (2, 7, 5) and te.size(fb.HH) == (1, 7, 5) and tm.length(fb.HH) == 70
    g1, _ = _rand_grid(rng, (9,), (False,))
    tem = fb.ModelTEM(g1)
    assert tem.eps_arr.shape == (9, 1, 1) and tem.cmp_s == (2,) and tem.cmp_e == (0,) and tem.cmp_m == (1,)
    for bad in (lambda: fb.ModelTE(g1), lambda: fb.ModelTEM(g2), lambda: fb.ModelFull(g2)):
        with pytest.raises(ValueError):
            bad()
    # DOF order: r = c + Kf (i + Nx j) (cmp-first) / i + Nx (j + Ny c)
    F = rng.standard_normal((7, 5, 2))
    v = fb.field_arr2vec(F)
    assert v[1 + 2 * (3 + 7 * 4)] == F[3, 4, 1] and np.array_equal(fb.field_vec2arr(v, (7, 5)), F)
    v = fb.field_arr2vec(F, order_cmpfirst=False)
    assert v[3 + 7 * (4 + 5 * 1)] == F[3, 4, 1] and np.array_equal(fb.field_vec2arr(v, (7, 5), False), F)
    # sources
    c = [rng.uniform(g2.bounds[0][w], g2.bounds[1][w]) for w in range(2)]
    assert not fb.PointSrc(c, [1, 1]).isfield_ortho_shp and fb.PointSrc(c, [1]).isfield_ortho_shp
    assert not fb.PointSrc([0, 0, 0], [0, 0, 1]).isfield_ortho_shp and fb.PlaneSrc([1, 0], 0.0, [1]).isfield_ortho_shp
    bft = (fb.EE, fb.HH)
    for mdl, Ke, Km in ((te, 2, 1), (tm, 1, 2)):
        fb.set_boundft(mdl, bft)
        pe, pm = [1.0, -2.0][:Ke], [0.5, 1.0][:Km]
        fb.add_srce(mdl, fb.PointSrc(c, pe, 0.3 + 0.1j))
        fb.add_srce(mdl, fb.PlaneSrc([0, 1], c[1], pe, 2.0))
        fb.add_srcm(mdl, fb.PointSrc(c, pm))
        je, jm = osrc.create_field_array(o2.N, Ke), osrc.create_field_array(o2.N, Km)
        osrc.add_src(je, og.EE, bft, o2, osrc.PointSrc(c, pe, 0.3 + 0.1j, isfield_ortho_shp=Ke == 1))
        osrc.add_src(je, og.EE, bft, o2, osrc.PlaneSrc([0, 1], c[1], pe, 2.0, isfield_ortho_shp=Ke == 1))
        osrc.add_src(jm, og.HH, bft, o2, osrc.PointSrc(c, pm, isfield_ortho_shp=Km == 1))
        assert je.any() and jm.any()
        assert np.allclose(mdl.je_arr, je, rtol=1e-14, atol=0) and np.allclose(mdl.jm_arr, jm, rtol=1e-14, atol=0)
        vje, vjm = fb.create_srcs(mdl)
        assert vje.shape == (mdl.length(fb.EE),) and vjm.shape == (mdl.length(fb.HH),)
    # descriptors: objects on reduced models are refused loudly, z-slabs too
    Ps, Cs = fb.create_paramops(te), fb.create_curls(te)
    assert Ps[0].arr is te.eps_arr and Cs[0].geom.cmp_m == (2,)
    with pytest.raises(ValueError):
        fb.create_A(fb.EE, 1.0, Ps, Cs, nranks=2, rank=0)
