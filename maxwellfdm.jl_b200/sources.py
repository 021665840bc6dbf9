"""Current sources of the product's host API: PointSrc, PlaneSrc and the two-point spreading weights.

Host-side mirror of reference src/source/{source,pointsrc,planesrc}.jl (same constructor argument
order, same ArgumentError conditions -> ValueError).  O(surface) integer/1-D work done once per source,
kept on the host exactly like the reference; the result (the J arrays) feeds fdfd_create_b on the GPU.
"""
import itertools

import numpy as np

from .grid import PRIM, DUAL, alter


def distweights(c, gt, bounds, l, dl, isbloch):
    """source.jl:208-306.  Returns (ind[2], wt[2]) with 0-based indices."""
    lneg, lpos = float(bounds[0]), float(bounds[1])
    c = float(c)
    if c < lneg or c > lpos:
        raise ValueError(f"c = {c} must be within bounds = {(lneg, lpos)}.")
    l = np.asarray(l, dtype=np.float64)
    dl = np.asarray(dl, dtype=np.float64)
    N = l.size
    if N <= 1 and not isbloch:
        raise ValueError(f"length(l) = {N} must be > 1 for symmetry boundary (= non-Bloch).")
    zeroing = gt == PRIM and not isbloch
    first, last = (1 if zeroing else 0), N - 1
    interior = l[first] <= c < l[last]
    if interior:
        i1 = int(np.searchsorted(l, c, side="right")) - 1
        i2 = i1 + 1
        r = (c - l[i1]) / (l[i2] - l[i1])
    else:
        if c < l[first]:
            i1, i2, bnd = first, N - 1, lneg
        else:
            i1, i2, bnd = last, 0, lpos
        if gt == PRIM:
            span = abs(l[i1] - bnd)
        elif isbloch:
            span = (lpos - lneg) - abs(l[i1] - l[i2])
        else:
            span = 2 * abs(l[i1] - bnd)
        r = abs(l[i1] - c) / span
    w1 = 1.0 / dl[i1]
    if interior or isbloch or zeroing:
        w1 *= 1.0 - r
    if c == l[i1] or (not interior and not isbloch):
        return (i1, i1), (w1, 0.0)
    w2 = 1.0 / dl[i2]
    if interior or isbloch:
        w2 *= r
    return (i1, i2), (w1, w2)


def _unit(v):
    v = np.asarray(v, dtype=np.float64)
    return v / np.sqrt((v * v).sum())


def _gt_cmp(src, nw, gt0):
    if src.isfield_ortho_shp:
        return tuple(gt0)
    return tuple(alter(g) if k == nw else g for k, g in enumerate(gt0))


class Source:
    pass


def isfield_ortho_shape(Kf, K):
    """MaxwellBase's default for `isfield˔shp` (pointsrc.jl:56,61; planesrc.jl:20,25; not in the reference tree): true
    when the field components span the orthogonal complement of the shape axes (TM: E_z over x-y; TE: H_z), false
    when they are the same space (3-D, TE's E, TM's H).  Override per source where this default does not fit."""
    return Kf + K == 3


class PointSrc(Source):
    """PointSrc(c, p, I∆r=1) - pointsrc.jl:51-62."""

    def __init__(self, c, p, Idr=1.0, isfield_ortho_shp=None):
        self.c, self.p, self.Idr = np.atleast_1d(np.asarray(c, float)), _unit(np.atleast_1d(p)), complex(Idr)
        self.isfield_ortho_shp = isfield_ortho_shape(self.p.size, self.c.size) if isfield_ortho_shp is None else bool(isfield_ortho_shp)

    def add(self, jarr, gt0, bounds, l, dl, isbloch):
        K = self.c.size
        for nw in range(self.p.size):
            gt = _gt_cmp(self, nw, gt0)
            iw = [distweights(self.c[u], gt[u], (bounds[0][u], bounds[1][u]), l[gt[u]][u], dl[gt[u]][u], isbloch[u])
                  for u in range(K)]
            amp = self.Idr * self.p[nw]
            for corner in itertools.product((0, 1), repeat=K):
                wt = np.prod([iw[u][1][corner[u]] for u in range(K)])
                jarr[tuple(iw[u][0][corner[u]] for u in range(K)) + (nw,)] += amp * wt


class PlaneSrc(Source):
    """PlaneSrc(n, c, p, J∆n=1) - planesrc.jl:14-32."""

    def __init__(self, n, c, p, Jdn=1.0, isfield_ortho_shp=None):
        n = np.atleast_1d(np.asarray(n, float))
        if np.count_nonzero(n) != 1:
            raise ValueError(f"n = {n} must be along Cartesian direction.")
        self.n, self.c, self.p, self.Jdn = _unit(n), float(c), _unit(np.atleast_1d(p)), complex(Jdn)
        self.isfield_ortho_shp = isfield_ortho_shape(self.p.size, self.n.size) if isfield_ortho_shp is None else bool(isfield_ortho_shp)

    def add(self, jarr, gt0, bounds, l, dl, isbloch):
        nn = int(np.argmax(self.n == 1))
        for nw in range(self.p.size):
            g = _gt_cmp(self, nw, gt0)[nn]
            ind, wt = distweights(self.c, g, (bounds[0][nn], bounds[1][nn]), l[g][nn], dl[g][nn], isbloch[nn])
            for k in (0, 1):
                sl = [slice(None)] * self.n.size + [nw]
                sl[nn] = ind[k]
                jarr[tuple(sl)] += self.Jdn * self.p[nw] * wt[k]
