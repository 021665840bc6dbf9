"""Host-side grid geometry and SC-PML stretch factors of the product (O(Nx+Ny+Nz) work).

Mirrors what the reference obtains from MaxwellBase: `Grid(lprim, isbloch)` with fields
N, L, l, ∆l, bounds, isbloch (used at reference src/model/model.jl:40,91,186-188 and
test/source.jl:129-130) and `create_stretched_∆l(ωpml, grid, Npml)` (model.jl:126).
Conventions (test/source.jl:6-15): l[PRIM] omits the +end ghost primal point, l[DUAL] omits the -end
ghost dual point, ∆l[PRIM] = diff(ghosted dual), ∆l[DUAL] = diff(ghosted primal).  0-based, PRIM=0, DUAL=1.
"""
import numpy as np

PRIM, DUAL = 0, 1
EE, HH = 0, 1


def ft2gt(ft, boundft):
    return PRIM if ft == boundft else DUAL


def alter(gt):
    return DUAL - gt


class Grid:
    def __init__(self, lprim, isbloch):
        self.lg_prim = tuple(np.array(a, dtype=np.float64) for a in lprim)     # ghosted primal
        self.isbloch = tuple(bool(b) for b in isbloch)
        self.N = tuple(a.size - 1 for a in self.lg_prim)
        self.L = tuple(float(a[-1] - a[0]) for a in self.lg_prim)
        self.bounds = (tuple(float(a[0]) for a in self.lg_prim), tuple(float(a[-1]) for a in self.lg_prim))
        lg_dual = []
        for a, L, bl in zip(self.lg_prim, self.L, self.isbloch):
            mid = (a[1:] + a[:-1]) / 2
            lg_dual.append(np.r_[mid[-1] - L if bl else 2 * a[0] - mid[0], mid])
        self.lg_dual = tuple(lg_dual)                                           # ghosted dual
        self.l = (tuple(a[:-1] for a in self.lg_prim), tuple(a[1:] for a in self.lg_dual))
        self.dl = (tuple(np.diff(a) for a in self.lg_dual), tuple(np.diff(a) for a in self.lg_prim))

    def __len__(self):
        return len(self.N)


class PMLParam:
    """Polynomial-graded SC-PML: s = kappa + sigma/(a + i w)  (exp(+iwt), model.jl:1-22)."""

    def __init__(self, m=4.0, R=np.exp(-16.0), kappa_max=1.0, a_max=0.0, m_a=4.0):
        self.m, self.R, self.kappa_max, self.a_max, self.m_a = m, R, kappa_max, a_max, m_a

    def s(self, depth_frac, dpml, w):
        x = depth_frac
        sigma = -(self.m + 1) * np.log(self.R) / (2 * dpml) * x ** self.m
        kappa = 1 + (self.kappa_max - 1) * x ** self.m
        a = self.a_max * (1 - x) ** self.m_a
        return kappa + sigma / (a + 1j * w)


def create_stretched_dl(wpml, grid, Npml, pml=None):
    """sdl[g][w][i] = s_w(l[g][w][i]) * dl[g][w][i]."""
    pml = pml or PMLParam()
    out = ([], [])
    for w, lp in enumerate(grid.lg_prim):
        lo, hi = lp[int(Npml[0][w])], lp[lp.size - 1 - int(Npml[1][w])]
        for g in (PRIM, DUAL):
            l = grid.l[g][w]
            s = np.ones(l.size, dtype=np.complex128)
            if lo > lp[0]:
                m = l < lo
                s[m] = pml.s((lo - l[m]) / (lo - lp[0]), lo - lp[0], wpml)
            if hi < lp[-1]:
                m = l > hi
                s[m] = pml.s((l[m] - hi) / (lp[-1] - hi), lp[-1] - hi, wpml)
            out[g].append(s * grid.dl[g][w])
    return tuple(out[0]), tuple(out[1])
