"""Pins oracle/source.py against the reference's own known-answer tests.

Restates reference test/source.jl: 'distweights' (:5-121), 'PlaneSrc' (:123-170),
'PointSrc' (:172-230).  Indices are 0-based here (Julia index - 1).
"""
import numpy as np
import pytest

from oracle.grid import Grid, PRIM, DUAL, EE, ft2gt
from oracle.source import distweights, PointSrc, PlaneSrc, add_src, create_field_array

nX, nY, nZ = 0, 1, 2


def _grid_1d():
    # test/source.jl:6-15
    dlg = np.arange(2, 23, 2)
    lp = np.cumsum(np.concatenate(([-1], dlg)))          # [-1,1,5,...,131]
    lprim_g = 0.5 * (lp[:-1] + lp[1:])                    # movingavg: [0,3,8,...,120]
    ldual_g = lp[:-1]                                     # [-1,1,...,109]
    assert list(lprim_g) == [0, 3, 8, 15, 24, 35, 48, 63, 80, 99, 120]
    domain = (lprim_g[0], lprim_g[-1])
    lprim, dlprim = lprim_g[:-1], np.diff(ldual_g)
    ldual, dldual = ldual_g[1:], np.diff(lprim_g)
    return domain, lprim, dlprim, ldual, dldual


def _chk(res, ind, wt):
    (i1, i2), (w1, w2) = res
    assert (i1, i2) == tuple(ind)
    assert np.allclose([w1, w2], wt, rtol=1e-14, atol=0)


def test_distweights_primal():  # test/source.jl:17-60
    domain, lprim, dlprim, _, _ = _grid_1d()
    dw = lambda c, b: distweights(c, PRIM, domain, lprim, dlprim, b)
    for b in (True, False):
        _chk(dw(8, b), (2, 2), [1 / dlprim[2], 0])
    r = (11 - 8) / (15 - 8)
    for b in (True, False):
        _chk(dw(11, b), (2, 3), [1 / dlprim[2] * (1 - r), 1 / dlprim[3] * r])
    for b in (True, False):
        _chk(dw(3, b), (1, 1), [1 / dlprim[1], 0])
    r = (3 - 1) / (3 - 0)
    _chk(dw(1, True), (0, 1), [1 / dlprim[0] * r, 1 / dlprim[1] * (1 - r)])
    _chk(dw(1, False), (1, 1), [1 / dlprim[1] * (1 - r), 0])
    _chk(dw(0, True), (0, 0), [1 / dlprim[0], 0])
    _chk(dw(0, False), (1, 1), [0, 0])
    for c in (-0.5, -1):
        for b in (True, False):
            with pytest.raises(ValueError):
                dw(c, b)
    for b in (True, False):
        _chk(dw(99, b), (9, 9), [1 / dlprim[9], 0])
    for c in (105, 120):
        r = (c - 99) / (120 - 99)
        _chk(dw(c, True), (9, 0), [1 / dlprim[9] * (1 - r), 1 / dlprim[0] * r])
        _chk(dw(c, False), (9, 9), [1 / dlprim[9] * (1 - r), 0])


def test_distweights_dual():  # test/source.jl:64-104
    domain, _, _, ldual, dldual = _grid_1d()
    dw = lambda c, b: distweights(c, DUAL, domain, ldual, dldual, b)
    for b in (True, False):
        _chk(dw(11, b), (2, 2), [1 / dldual[2], 0])
    r = (15 - 11) / (19 - 11)
    for b in (True, False):
        _chk(dw(15, b), (2, 3), [1 / dldual[2] * (1 - r), 1 / dldual[3] * r])
    for b in (True, False):
        _chk(dw(1, b), (0, 0), [1 / dldual[0], 0])
    for c in (0.5, 0):
        r = (1 - c) / ((1 - 0) + (120 - 109))
        _chk(dw(c, True), (0, 9), [1 / dldual[0] * (1 - r), 1 / dldual[9] * r])
        _chk(dw(c, False), (0, 0), [1 / dldual[0], 0])
    for b in (True, False):
        with pytest.raises(ValueError):
            dw(-0.5, b)
    for b in (True, False):
        _chk(dw(109, b), (9, 9), [1 / dldual[9], 0])
    for c in (115, 120):
        r = (c - 109) / ((120 - 109) + (1 - 0))
        _chk(dw(c, True), (9, 0), [1 / dldual[9] * (1 - r), 1 / dldual[0] * r])
        _chk(dw(c, False), (9, 9), [1 / dldual[9], 0])


def test_distweights_N1():  # test/source.jl:106-120
    rng = np.random.default_rng(1)
    domain = (0, 2)
    for c in 2 * rng.random(8):
        ind, wt = distweights(c, PRIM, domain, [0], [9], True)
        assert ind == (0, 0) and np.isclose(sum(wt), 1 / 9)
        ind, wt = distweights(c, DUAL, domain, [1], [10], True)
        assert ind == (0, 0) and np.isclose(sum(wt), 1 / 10)
    with pytest.raises(ValueError):
        distweights(1.0, PRIM, domain, [0], [9], False)


def _grid3(lprim):
    return Grid(lprim, (True, True, True))


def test_planesrc():  # test/source.jl:123-170
    src = PlaneSrc([0, 0, 1], 0, [1, 0, 0])
    boundft = (EE, EE, EE)
    g3 = _grid3((np.arange(-10, 11.0),) * 3)
    j3d = create_field_array(g3.N)
    add_src(j3d, EE, boundft, g3, src)
    assert np.abs(j3d).max() == 1.0
    assert not j3d[..., nY].any() and not j3d[..., nZ].any()

    fine = np.arange(-10, 10.25, 0.5)
    g3f = _grid3((fine,) * 3)
    j3f = create_field_array(g3f.N)
    add_src(j3f, EE, boundft, g3f, src)
    assert np.isclose(j3d[0, :, :, nX].sum() * 1.0, j3f[0, :, :, nX].sum() * 0.25)
    assert np.abs(j3f).max() == 2.0

    rng = np.random.default_rng(2)
    zprim = np.sort(rng.random(21)) * 20
    zprim -= zprim.mean()
    g3n = _grid3((np.arange(-10, 11.0), np.arange(-10, 11.0), zprim))
    j3n = create_field_array(g3n.N)
    add_src(j3n, EE, boundft, g3n, src)
    dy, dz = g3n.dl[PRIM][nY], g3n.dl[PRIM][nZ]
    assert np.isclose(j3d[0, :, :, nX].sum(), (j3n[0, :, :, nX] * np.outer(dy, dz)).sum())
    with pytest.raises(ValueError):
        PlaneSrc([1, 1, 0], 0, [1, 0, 0])


def test_pointsrc():  # test/source.jl:172-230
    src = PointSrc([0.7, 0.7, 0.7], [1, 1, 1])
    boundft = (EE, EE, EE)
    for lp, dv in ((np.arange(-10, 11.0), 1.0), (np.arange(-10, 10.25, 0.5), 0.125)):
        g3 = _grid3((lp,) * 3)
        j3d = create_field_array(g3.N)
        add_src(j3d, EE, boundft, g3, src)
        for c in range(3):
            assert np.count_nonzero(j3d[..., c]) == 8
            # reference asserts exact ==; the uniform-grid sums are exact in binary here too
            assert j3d[..., c].sum() * dv == src.Idr * src.p[c]
    g3 = _grid3((np.arange(-10, 11.0),) * 3)
    j3d = create_field_array(g3.N)
    add_src(j3d, EE, boundft, g3, src)
    rng = np.random.default_rng(3)
    lps = []
    for _ in range(3):
        a = np.sort(rng.random(21)) * 20
        a -= a.mean() - 0.7
        lps.append(a)
    g3n = _grid3(tuple(lps))
    j3n = create_field_array(g3n.N)
    add_src(j3n, EE, boundft, g3n, src)
    dx, dy, dzd = g3n.dl[PRIM][nX], g3n.dl[PRIM][nY], g3n.dl[DUAL][nZ]
    vol = dx[:, None, None] * dy[None, :, None] * dzd[None, None, :]
    assert np.isclose(j3d[..., nZ].sum() * 1.0, (j3n[..., nZ] * vol).sum())
