"""Current-source discretisation - oracle restatement (TEST INFRASTRUCTURE ONLY).

Restates, with 0-based indices:
  distweights        reference src/source/source.jl:208-306
  PointSrc.add_src!  reference src/source/pointsrc.jl:51-103
  PlaneSrc.add_src!  reference src/source/planesrc.jl:14-80
  add_srce!/add_srcm! argument wiring  src/model/model.jl:184-200
This part of the oracle IS pinned: tests/test_oracle_source.py restates the reference's own
known-answer table test/source.jl:5-230.
"""
from __future__ import annotations

import numpy as np

from .grid import PRIM, DUAL, EE, HH, Grid, ft2gt


def alter(gt):
    return DUAL if gt == PRIM else PRIM


def gt_w(nw, gt0):
    """MaxwellBase.gt_w: grid types of the w-component of a field whose voxel corners have grid
    type gt0 - the component is shifted by half a cell along its own axis (pointsrc.jl:75)."""
    return tuple(alter(g) if k == nw else g for k, g in enumerate(gt0))


def distweights(c, gt, bounds, l, dl, isbloch):
    """Two-point current-spreading weights (source.jl:208-306).  Returns ((ind1, ind2), (wt1, wt2))
    with 0-based indices; raises ValueError where the reference throws ArgumentError."""
    c = float(c)
    l = np.asarray(l, dtype=np.float64)
    dl = np.asarray(dl, dtype=np.float64)
    if not (bounds[0] <= c <= bounds[1]):
        raise ValueError(f"c = {c} must be within bounds = {tuple(bounds)}.")
    L = bounds[1] - bounds[0]
    N = len(l)
    if not (N > 1 or isbloch):
        raise ValueError(f"length(l) = {N} must be > 1 for symmetry boundary (= non-Bloch).")

    zeroing_bc = (gt == PRIM) and (not isbloch)
    indn = 1 if zeroing_bc else 0
    indp = N - 1

    bc_noeff = l[indn] <= c < l[indp]
    if bc_noeff:
        ind1 = int(np.flatnonzero(l - c <= 0)[-1])
        ind2 = ind1 + 1
        dc1 = c - l[ind1]
        dlc = l[ind2] - l[ind1]
    else:
        if c < l[indn]:
            ind1, ind2, bnd = indn, N - 1, bounds[0]
        else:
            ind1, ind2, bnd = indp, 0, bounds[1]
        dc1 = abs(l[ind1] - c)
        if gt == PRIM:
            dlc = abs(l[ind1] - bnd)
        else:
            dlc = (L - abs(l[ind1] - l[ind2])) if isbloch else 2.0 * abs(l[ind1] - bnd)

    r = dc1 / dlc
    wt1 = 1.0 / dl[ind1]
    if bc_noeff or isbloch or zeroing_bc:
        wt1 *= 1.0 - r
    if c == l[ind1] or ((not bc_noeff) and (not isbloch)):
        ind2 = ind1
        wt2 = 0.0
    else:
        wt2 = 1.0 / dl[ind2]
        if bc_noeff or isbloch:
            wt2 *= r
    return (ind1, ind2), (wt1, wt2)


def _normalize(p):
    p = np.asarray(p, dtype=np.float64)
    return p / np.linalg.norm(p)


class PointSrc:
    """PointSrc(c, p, Idr=1) (pointsrc.jl:51-62); p is normalised like the reference ctor."""

    def __init__(self, c, p, Idr=1.0, isfield_ortho_shp=False):
        self.c = np.asarray(c, dtype=np.float64)
        self.p = _normalize(p)
        self.Idr = complex(Idr)
        self.isfield_ortho_shp = isfield_ortho_shp

    def add_to(self, jarr, gt0, bounds, l, dl, isbloch):
        """add_src!(jKd, gt0, bounds, l, dl, isbloch, src::PointSrc) (pointsrc.jl:64-103)."""
        K = len(self.c)
        for nw in range(len(self.p)):
            gt_cmp = tuple(gt0) if self.isfield_ortho_shp else gt_w(nw, gt0)
            inds, wts = [], []
            for nu in range(K):
                g = gt_cmp[nu]
                ind_u, wt_u = distweights(self.c[nu], g, (bounds[0][nu], bounds[1][nu]),
                                          l[g][nu], dl[g][nu], isbloch[nu])
                inds.append(ind_u)
                wts.append(wt_u)
            Iw = self.Idr * self.p[nw]
            for corner in np.ndindex(*([2] * K)):
                wt = 1.0
                for nu in range(K):
                    wt = wt * wts[nu][corner[nu]]
                idx = tuple(inds[nu][corner[nu]] for nu in range(K))
                jarr[idx + (nw,)] += Iw * wt


class PlaneSrc:
    """PlaneSrc(n, c, p, Jdn=1) (planesrc.jl:14-32); n must be a Cartesian direction."""

    def __init__(self, n, c, p, Jdn=1.0, isfield_ortho_shp=False):
        n = np.asarray(n, dtype=np.float64)
        if np.count_nonzero(n) != 1:
            raise ValueError(f"n = {n} must be along Cartesian direction.")
        self.n = _normalize(n)
        self.c = float(c)
        self.p = _normalize(p)
        self.Jdn = complex(Jdn)
        self.isfield_ortho_shp = isfield_ortho_shp

    def add_to(self, jarr, gt0, bounds, l, dl, isbloch):
        """add_src!(jKd, gt0, bounds, l, dl, isbloch, src::PlaneSrc) (planesrc.jl:34-80)."""
        nn = int(np.flatnonzero(self.n == 1)[0])
        K = len(self.n)
        for nw in range(len(self.p)):
            gt_cmp = tuple(gt0) if self.isfield_ortho_shp else gt_w(nw, gt0)
            g = gt_cmp[nn]
            ind, wt = distweights(self.c, g, (bounds[0][nn], bounds[1][nn]), l[g][nn], dl[g][nn], isbloch[nn])
            Jw = self.Jdn * self.p[nw]
            for k in range(2):
                sl = [slice(None)] * K
                sl[nn] = ind[k]
                jarr[tuple(sl) + (nw,)] += Jw * wt[k]


def add_src(jarr, ft, boundft, grid: Grid, src):
    """add_srce!/add_srcm! (model.jl:184-200): ft = EE for electric, HH for magnetic current."""
    gt0 = tuple(ft2gt(ft, b) for b in boundft)
    src.add_to(jarr, gt0, grid.bounds, grid.l, grid.dl, grid.isbloch)


def create_field_array(N, ncmp=3):
    """MaxwellBase.create_field_array (call sites model.jl:55-56): zeros of shape (N..., ncmp)."""
    return np.zeros(tuple(N) + (ncmp,), dtype=np.complex128)


def create_param_array(N, ncmp=3):
    """MaxwellBase.create_param_array (call sites model.jl:51-52): zeros (N..., ncmp, ncmp)."""
    return np.zeros(tuple(N) + (ncmp, ncmp), dtype=np.complex128)
