"""Matrix-free numpy restatement of y = A x (TEST INFRASTRUCTURE ONLY).

An implementation of the same operator as oracle/operators.py that shares NO code with it
(array slicing / np.roll on (Nx,Ny,Nz,3) arrays instead of sparse matrices), so that property
P6 of SURVEY.md §8c (CSC builder vs matrix-free apply agree to 1e-13) is a real cross-check.
It also serves grids whose CSC does not fit (configs 4-5), optionally on z-sub-slabs.

Operator: reference create_A (src/model/model.jl:225-246) with the create_curls /
create_paramops argument polarity of model.jl:141-175; out-of-tree pieces per SURVEY App. A.
"""
from __future__ import annotations

import numpy as np

from .grid import EE, HH


def _bshape(a, w):
    shp = [1, 1, 1]
    shp[w] = len(a)
    return np.asarray(a, np.complex128).reshape(shp)


def _take(F, w, i):
    sl = [slice(None)] * 3
    sl[w] = slice(i, i + 1)
    return tuple(sl)


def diff(F, w, isfwd, dinv, isbloch, ph):
    """(D f)[i] along axis w of a scalar 3-D array."""
    N = F.shape[w]
    if isfwd:
        nxt = np.roll(F, -1, axis=w).copy()
        nxt[_take(F, w, N - 1)] *= ph if isbloch else 0.0
        cur = F.copy()
        if not isbloch:
            cur[_take(F, w, 0)] = 0.0
        return _bshape(dinv, w) * (nxt - cur)
    prv = np.roll(F, +1, axis=w).copy()
    cur = F.copy()
    if isbloch:
        prv[_take(F, w, 0)] /= ph
    else:
        prv[_take(F, w, 0)] = 0.0
        cur[_take(F, w, 0)] = 0.0
    return _bshape(dinv, w) * (cur - prv)


def mean(F, w, isfwd, dl, dlo_inv, isbloch, ph):
    """(M f)[i] along axis w; dl=None and dlo_inv=None => unweighted."""
    N = F.shape[w]
    dl = np.ones(N, np.complex128) if dl is None else np.asarray(dl, np.complex128)
    dlo_inv = np.ones(N, np.complex128) if dlo_inv is None else np.asarray(dlo_inv, np.complex128)
    G = _bshape(dl, w) * F
    if isfwd:
        nxt = np.roll(G, -1, axis=w).copy()
        nxt[_take(F, w, N - 1)] *= ph if isbloch else 0.0
        cur = G.copy()
        if not isbloch:
            cur[_take(F, w, 0)] = 0.0
        return 0.5 * _bshape(dlo_inv, w) * (cur + nxt)
    prv = np.roll(G, +1, axis=w).copy()
    cur = G.copy()
    if isbloch:
        prv[_take(F, w, 0)] /= ph
    else:
        prv[_take(F, w, 0)] = 0.0
        cur[_take(F, w, 0)] *= 2.0
    return 0.5 * _bshape(dlo_inv, w) * (cur + prv)


def curl(F, isfwd, dinv, isbloch, ph):
    """F[...,3] -> curl, component v = d_{v+1} F_{v+2} - d_{v+2} F_{v+1} (cyclic)."""
    out = np.empty_like(F)
    for v in range(3):
        w1, w2 = (v + 1) % 3, (v + 2) % 3
        out[..., v] = (diff(F[..., w2], w1, isfwd[w1], dinv[w1], isbloch[w1], ph[w1])
                       - diff(F[..., w1], w2, isfwd[w2], dinv[w2], isbloch[w2], ph[w2]))
    return out


def paramop(F, prm, isfwd_in, dl, dlo_inv, isbloch, ph, weighted_out=False, diag_only=False):
    """(P F)_v = p_vv F_v + sum_{u!=v} Mout_v[ p_vu * Min_u F_u ]   (SURVEY A.5)."""
    out = np.empty_like(F)
    for v in range(3):
        out[..., v] = prm[..., v, v] * F[..., v]
    if diag_only:
        return out
    avg = [mean(F[..., u], u, isfwd_in[u], dl[u], dlo_inv[u], isbloch[u], ph[u]) for u in range(3)]
    for v in range(3):
        g = np.zeros(F.shape[:3], np.complex128)
        for u in range(3):
            if u != v:
                g += prm[..., v, u] * avg[u]
        if weighted_out:
            out[..., v] += mean(g, v, not isfwd_in[v], 1.0 / np.asarray(dlo_inv[v]),
                                1.0 / np.asarray(dl[v]), isbloch[v], ph[v])
        else:
            out[..., v] += mean(g, v, not isfwd_in[v], None, None, isbloch[v], ph[v])
    return out


class MatFreeOperator:
    """y = A x for ft in {EE, HH}; x, y are DOF vectors in the reference's ordering
    (model.jl:75-83) or (Nx,Ny,Nz,3) arrays via apply_arr()."""

    def __init__(self, ft, omega, eps, mu, sdl_e, sdl_m, boundft, isbloch, e_mikL,
                 order_cmpfirst=True, weighted_out=False):
        self.ft, self.omega = ft, omega
        self.eps = np.asarray(eps, np.complex128)
        self.N = self.eps.shape[:3]
        self.mu = None if mu is None else np.asarray(mu, np.complex128)
        self.sdl_e = [np.asarray(a, np.complex128) for a in sdl_e]
        self.sdl_m = [np.asarray(a, np.complex128) for a in sdl_m]
        self.boundft = tuple(boundft)
        self.isbloch = tuple(bool(b) for b in isbloch)
        self.ph = np.asarray(e_mikL, np.complex128)
        self.order_cmpfirst = order_cmpfirst
        self.weighted_out = weighted_out

        def offdiag_zero(p):
            if p is None:
                return True
            o = p.copy()
            for v in range(3):
                o[..., v, v] = 0
            return not o.any()
        self.eps_diag_only = offdiag_zero(self.eps)
        self.mu_diag_only = offdiag_zero(self.mu)

    # -- pieces -------------------------------------------------------------------------
    def Ce(self, E):
        isfwd = [b == EE for b in self.boundft]
        return curl(E, isfwd, [1.0 / a for a in self.sdl_m], self.isbloch, self.ph)

    def Cm(self, H):
        isfwd = [b == HH for b in self.boundft]
        return curl(H, isfwd, [1.0 / a for a in self.sdl_e], self.isbloch, self.ph)

    def Peps(self, E):
        isfwd_in = [b != EE for b in self.boundft]
        return paramop(E, self.eps, isfwd_in, self.sdl_m, [1.0 / a for a in self.sdl_e],
                       self.isbloch, self.ph, self.weighted_out, self.eps_diag_only)

    def Pmu(self, H):
        if self.mu is None:
            return H.copy()
        isfwd_in = [b != HH for b in self.boundft]
        return paramop(H, self.mu, isfwd_in, self.sdl_e, [1.0 / a for a in self.sdl_m],
                       self.isbloch, self.ph, self.weighted_out, self.mu_diag_only)

    def _mu_inv(self, H):
        if self.mu is None:
            return H
        if not self.mu_diag_only:
            raise ValueError("Pmu must be diagonal (reference model.jl:236)")
        return H / np.stack([self.mu[..., v, v] for v in range(3)], axis=-1)

    def _eps_inv(self, E):
        if not self.eps_diag_only:
            raise ValueError("Peps must be diagonal (reference model.jl:239)")
        return E / np.stack([self.eps[..., v, v] for v in range(3)], axis=-1)

    # -- operator -----------------------------------------------------------------------
    def apply_arr(self, X):
        X = np.asarray(X, np.complex128)
        if self.ft == EE:
            Y = self.Cm(self._mu_inv(self.Ce(X)))
            if self.omega != 0:
                Y = Y - self.omega ** 2 * self.Peps(X)
        elif self.ft == HH:
            Y = self.Ce(self._eps_inv(self.Cm(X)))
            if self.omega != 0:
                Y = Y - self.omega ** 2 * self.Pmu(X)
        else:
            raise ValueError(f"ft = {self.ft} is unsupported.")
        return Y

    def vec2arr(self, x):
        Nx, Ny, Nz = self.N
        if self.order_cmpfirst:
            return np.asarray(x).reshape(Nz, Ny, Nx, 3).transpose(2, 1, 0, 3)
        return np.asarray(x).reshape(3, Nz, Ny, Nx).transpose(3, 2, 1, 0)

    def arr2vec(self, F):
        if self.order_cmpfirst:
            return np.ascontiguousarray(F.transpose(2, 1, 0, 3)).ravel()
        return np.ascontiguousarray(F.transpose(3, 2, 1, 0)).ravel()

    def apply(self, x):
        return self.arr2vec(self.apply_arr(self.vec2arr(x)))

    __call__ = apply
