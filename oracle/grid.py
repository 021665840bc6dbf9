"""Staggered (Yee) grid geometry and SC-PML stretch factors - oracle restatement.

Follows:
  * grid conventions evidenced in-tree at reference test/source.jl:6-15 and
    src/source/source.jl:128-131 (l[PRIM] drops the +end ghost primal point, l[DUAL] drops
    the -end ghost dual point, dl[PRIM] = diff(ghosted dual), dl[DUAL] = diff(ghosted primal));
  * the consumer of the stretched cell sizes, src/model/model.jl:122-139
    (create_stretched_dls: which array goes to E-planes / H-planes);
  * MaxwellBase.create_stretched_dl (NOT in tree; SURVEY.md Appendix A.1 restatement).
All indices here are 0-based; PRIM = 0, DUAL = 1 (Julia nPR = 1, nDL = 2).
"""
from __future__ import annotations

from dataclasses import dataclass, field
import numpy as np

PRIM, DUAL = 0, 1
EE, HH = 0, 1  # FieldType enum of the reference (EE: E-field, HH: H-field)


def ft2gt(ft: int, boundft: int) -> int:
    """MaxwellBase.ft2gt (call sites model.jl:130-131): PRIM iff ft == boundft."""
    return PRIM if ft == boundft else DUAL


class Grid:
    """Grid(lprim_g, isbloch): lprim_g[w] are the N_w+1 primal plane positions INCLUDING the
    positive-end ghost point (as passed to the reference's Grid ctor, test/source.jl:129-130)."""

    def __init__(self, lprim_g, isbloch):
        self.lprim_g = tuple(np.asarray(a, dtype=np.float64) for a in lprim_g)
        self.K = len(self.lprim_g)
        self.isbloch = tuple(bool(b) for b in isbloch)
        assert len(self.isbloch) == self.K
        self.N = tuple(len(a) - 1 for a in self.lprim_g)
        self.L = tuple(float(a[-1] - a[0]) for a in self.lprim_g)
        self.bounds = (tuple(float(a[0]) for a in self.lprim_g),
                       tuple(float(a[-1]) for a in self.lprim_g))
        ldual_g = []
        for w, lp in enumerate(self.lprim_g):
            ld = 0.5 * (lp[:-1] + lp[1:])  # dual points = midpoints of primal
            ghost = ld[-1] - self.L[w] if self.isbloch[w] else 2.0 * lp[0] - ld[0]
            ldual_g.append(np.concatenate(([ghost], ld)))
        self.ldual_g = tuple(ldual_g)
        # l[g][w], dl[g][w]: N_w entries each, no ghost points
        self.l = (tuple(a[:-1].copy() for a in self.lprim_g), tuple(a[1:].copy() for a in self.ldual_g))
        self.dl = (tuple(np.diff(a) for a in self.ldual_g), tuple(np.diff(a) for a in self.lprim_g))


@dataclass
class PMLParam:
    """SC-PML profile constants (defaults are a recollection of MaxwellBase and are
    configurable because parity tests feed identical 1-D arrays to oracle and GPU)."""
    m: float = 4.0
    R: float = float(np.exp(-16.0))
    kappa_max: float = 1.0
    a_max: float = 0.0
    m_a: float = 4.0


def _s_factor(omega_pml, l, lneg, lpos, lpml_neg, lpml_pos, prm: PMLParam):
    """s_w(l) = kappa + sigma / (a + i*omega) inside the PML, 1 outside (exp(+i w t) convention,
    reference model.jl:1-22)."""
    s = np.ones_like(l, dtype=np.complex128)
    for side in (0, 1):
        if side == 0:
            dpml = lpml_neg - lneg
            mask = l < lpml_neg
            d = lpml_neg - l
        else:
            dpml = lpos - lpml_pos
            mask = l > lpml_pos
            d = l - lpml_pos
        if dpml <= 0 or not mask.any():
            continue
        x = d[mask] / dpml
        sigma_max = -(prm.m + 1.0) * np.log(prm.R) / (2.0 * dpml)
        sigma = sigma_max * x ** prm.m
        kappa = 1.0 + (prm.kappa_max - 1.0) * x ** prm.m
        a = prm.a_max * (1.0 - x) ** prm.m_a
        s[mask] = kappa + sigma / (a + 1j * omega_pml)
    return s


def create_stretched_dl(omega_pml, grid: Grid, Npml, prm: PMLParam | None = None):
    """MaxwellBase.create_stretched_dl(wpml, grid, Npml) (call site model.jl:126).
    Npml = (Npml_neg[K], Npml_pos[K]) in cells.  Returns sdl[g][w] complex arrays."""
    prm = prm or PMLParam()
    out = ([], [])
    for w in range(grid.K):
        lp = grid.lprim_g[w]
        nneg, npos = int(Npml[0][w]), int(Npml[1][w])
        lpml_neg, lpml_pos = lp[nneg], lp[len(lp) - 1 - npos]
        for g in (PRIM, DUAL):
            s = _s_factor(omega_pml, grid.l[g][w], lp[0], lp[-1], lpml_neg, lpml_pos, prm)
            out[g].append(s * grid.dl[g][w])
    return (tuple(out[0]), tuple(out[1]))


def create_stretched_dls(omega_pml, grid: Grid, Npml, boundft=(EE, EE, EE), prm=None):
    """Reference create_stretched_dls (model.jl:122-139): returns
    (sdl_e, sdl_m, sdl_e_inv, sdl_m_inv), each a K-tuple of 1-D complex arrays, where
    sdl_e are centred at E-field plane locations (g_e = ft2gt.(EE, boundft)) and sdl_m at
    H-field plane locations."""
    sdl = create_stretched_dl(omega_pml, grid, Npml, prm)
    ge = [ft2gt(EE, b) for b in boundft]
    gm = [ft2gt(HH, b) for b in boundft]
    sdl_e = tuple(sdl[ge[w]][w] for w in range(grid.K))
    sdl_m = tuple(sdl[gm[w]][w] for w in range(grid.K))
    return sdl_e, sdl_m, tuple(1.0 / a for a in sdl_e), tuple(1.0 / a for a in sdl_m)


def create_e_mikL(kbloch, grid: Grid):
    """Reference create_e^{-ikL} (model.jl:91): exp.(-im .* kbloch .* L)."""
    return np.exp(-1j * np.asarray(kbloch, dtype=np.float64) * np.asarray(grid.L))
