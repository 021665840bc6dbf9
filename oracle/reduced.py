"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's 2-D / 1-D operators (ModelTE te.jl:4-14, ModelTM
tm.jl:4-14, ModelTEM tem.jl:4-13), assembled on the K-DIMENSIONAL grid from the scalar difference / averaging
operators of oracle/operators.py (create_d, create_m with a K-tuple N): this is what the reference's
create_curl(...; cmp_shp, cmp_out, cmp_in) and create_paramop produce for K < 3 (call sites model.jl:152-155,171-172).
It never touches a 3-D grid, so it is independent of the embedding the product uses (maxwellfdm.jl_b200/reduced.py).
parity unpinned, as for the 3-D operator: the reference's tests hold no vectors for create_curls / create_A.
"""
import numpy as np
import scipy.sparse as sp

from .grid import EE, HH
from .operators import create_d, create_m, _levi_civita

TE = dict(cmp_s=(0, 1), cmp_e=(0, 1), cmp_m=(2,))
TM = dict(cmp_s=(0, 1), cmp_e=(2,), cmp_m=(0, 1))
TEM = dict(cmp_s=(2,), cmp_e=(0,), cmp_m=(1,))


def _dof(cellmat_blocks, Kout, Kin, M, order_cmpfirst):
    """assemble a (Kout*M) x (Kin*M) matrix from blocks[(v,u)] (M x M), DOF order of model.jl:75-83"""
    rows = []
    for v in range(Kout):
        rows.append([cellmat_blocks.get((v, u)) for u in range(Kin)])
    A = sp.bmat(rows, format="csr")           # component-major (cmp-last order)
    if not order_cmpfirst:
        return A.tocsc()
    pr = (np.arange(Kout * M) % Kout) * M + np.arange(Kout * M) // Kout     # cmp-first row r -> cmp-last index
    pc = (np.arange(Kin * M) % Kin) * M + np.arange(Kin * M) // Kin
    return A[pr][:, pc].tocsc()


def create_curl(isfwd, dl_inv, isbloch, e_mikL, cmp_s, cmp_out, cmp_in, order_cmpfirst=True):
    """block (v,u) = eps_{v w u} D_w for the axis w completing (v,u), zero when w is not a grid axis"""
    N = tuple(len(a) for a in dl_inv)
    M = int(np.prod(N))
    blocks = {}
    for iv, v in enumerate(cmp_out):
        for iu, u in enumerate(cmp_in):
            if u == v:
                continue
            w = 3 - u - v
            if w not in cmp_s:
                continue
            k = cmp_s.index(w)
            D = create_d(k, bool(isfwd[k]), N, dl_inv[k], bool(isbloch[k]), e_mikL[k]).to_scipy()
            blocks[(iv, iu)] = _levi_civita(v, w, u) * D
    for iv in range(len(cmp_out)):
        for iu in range(len(cmp_in)):
            blocks.setdefault((iv, iu), sp.csr_matrix((M, M), dtype=complex))
    return _dof(blocks, len(cmp_out), len(cmp_in), M, order_cmpfirst)


def create_paramop(param, cmps, cmp_s, isfwd_in, dl, dl_out_inv, isbloch, e_mikL, order_cmpfirst=True):
    """Kf = 1: diag(param) (model.jl:152,154 first branch); Kf >= 2: diagonal entries plus, for u != v,
    Mout_v [ p_vu .* (Min_u F_u) ] with the averaging arguments of model.jl:153,155"""
    param = np.asarray(param, np.complex128)
    Kf = len(cmps)
    N = param.shape[:-2]
    M = int(np.prod(N))
    blocks = {}
    for i in range(Kf):
        blocks[(i, i)] = sp.diags(param[..., i, i].ravel(order="F")).tocsr()
    if Kf > 1:
        for i, v in enumerate(cmps):
            for j, u in enumerate(cmps):
                if i == j or not param[..., i, j].any():
                    continue
                ku, kv = cmp_s.index(u), cmp_s.index(v)
                Min = create_m(ku, bool(isfwd_in[ku]), N, dl[ku], dl_out_inv[ku], bool(isbloch[ku]), e_mikL[ku]).to_scipy()
                Mout = create_m(kv, not bool(isfwd_in[kv]), N, None, None, bool(isbloch[kv]), e_mikL[kv]).to_scipy()
                blocks[(i, j)] = (Mout @ sp.diags(param[..., i, j].ravel(order="F")) @ Min).tocsr()
    for i in range(Kf):
        for j in range(Kf):
            blocks.setdefault((i, j), sp.csr_matrix((M, M), dtype=complex))
    return _dof(blocks, Kf, Kf, M, order_cmpfirst)


class ReducedSystem:
    """Pe, Pm, Ce, Cm and the compositions of model.jl:225-284 for a K-dimensional model."""

    def __init__(self, kind, eps, mu, sdl_e, sdl_m, boundft, isbloch, e_mikL, order_cmpfirst=True):
        cs, ce, cm = kind["cmp_s"], kind["cmp_e"], kind["cmp_m"]
        sei, smi = [1 / np.asarray(a) for a in sdl_e], [1 / np.asarray(a) for a in sdl_m]
        self.Ce = create_curl([b == EE for b in boundft], smi, isbloch, e_mikL, cs, cm, ce, order_cmpfirst)
        self.Cm = create_curl([b == HH for b in boundft], sei, isbloch, e_mikL, cs, ce, cm, order_cmpfirst)
        self.Pe = create_paramop(eps, ce, cs, [b != EE for b in boundft], sdl_m, sei, isbloch, e_mikL, order_cmpfirst)
        self.Pm = create_paramop(mu, cm, cs, [b != HH for b in boundft], sdl_e, smi, isbloch, e_mikL, order_cmpfirst)

    def A(self, ft, omega):
        if ft == EE:
            return (self.Cm @ sp.diags(1 / self.Pm.diagonal()) @ self.Ce - omega ** 2 * self.Pe).tocsc()
        return (self.Ce @ sp.diags(1 / self.Pe.diagonal()) @ self.Cm - omega ** 2 * self.Pm).tocsc()

    def b(self, ft, omega, je, jm):
        if ft == EE:
            return -(self.Cm @ (jm / self.Pm.diagonal())) - 1j * omega * je
        return self.Ce @ (je / self.Pe.diagonal()) - 1j * omega * jm

    def h_from_e(self, e, omega, jm):
        return (1j / omega) * ((self.Ce @ e + jm) / self.Pm.diagonal())

    def e_from_h(self, h, omega, je):
        return (-1j / omega) * ((self.Cm @ h - je) / self.Pe.diagonal())


# ---- the same operators with the storage structure Julia would give them (for the index-pattern export) ----------------
def julia_csc(kind, ft, omega, eps, mu, sdl_e, sdl_m, boundft, isbloch, e_mikL, order_cmpfirst=True):
    """A of a K-dimensional model as a Csc with Julia's structure: every factor from `sparse(I,J,V)` triplets
    (duplicates summed, explicit zeros kept), products and the mass term with the structural rules of
    oracle/operators.py (spgemm, diag_ldiv, sub_scaled) - the composition of model.jl:225-246."""
    from .operators import (create_d_info, create_m, coo_to_csc, dof_index, spgemm, Csc, create_A as compose)
    cs, ce, cm = kind["cmp_s"], kind["cmp_e"], kind["cmp_m"]
    N = tuple(len(a) for a in sdl_e)
    M = int(np.prod(N))
    sei, smi = [1 / np.asarray(a) for a in sdl_e], [1 / np.asarray(a) for a in sdl_m]

    def curl(isfwd, dl_inv, cout, cin):
        Is, Js, Vs = [np.zeros(0, np.int64)], [np.zeros(0, np.int64)], [np.zeros(0, complex)]
        for iv, v in enumerate(cout):
            for iu, u in enumerate(cin):
                w = 3 - u - v
                if u == v or w not in cs:
                    continue
                k = cs.index(w)
                I, J, V = create_d_info(k, bool(isfwd[k]), N, dl_inv[k], bool(isbloch[k]), e_mikL[k])
                Is.append(dof_index(I, iv, M, len(cout), order_cmpfirst))
                Js.append(dof_index(J, iu, M, len(cin), order_cmpfirst))
                Vs.append(_levi_civita(v, w, u) * V)
        return coo_to_csc(np.concatenate(Is), np.concatenate(Js), np.concatenate(Vs), (len(cout) * M, len(cin) * M))

    def paramop(param, cmps, isfwd_in, dl, dlo_inv):
        param = np.asarray(param, np.complex128)
        Kf = len(cmps)
        cell = np.arange(M, dtype=np.int64)
        Is, Js, Vs = [], [], []
        for i in range(Kf):
            Is.append(dof_index(cell, i, M, Kf, order_cmpfirst)); Js.append(Is[-1]); Vs.append(param[..., i, i].ravel(order="F"))
        off = any(param[..., i, j].any() for i in range(Kf) for j in range(Kf) if i != j)
        if off:
            for i, v in enumerate(cmps):
                for j, u in enumerate(cmps):
                    if i == j:
                        continue
                    ku, kv = cs.index(u), cs.index(v)
                    Min = create_m(ku, bool(isfwd_in[ku]), N, dl[ku], dlo_inv[ku], bool(isbloch[ku]), e_mikL[ku])
                    Mout = create_m(kv, not bool(isfwd_in[kv]), N, None, None, bool(isbloch[kv]), e_mikL[kv])
                    pvu = param[..., i, j].ravel(order="F")
                    blk = spgemm(Mout, Csc(Min.shape, Min.colptr, Min.rowval, Min.nzval * pvu[Min.rowval]))
                    Is.append(dof_index(blk.rowval, i, M, Kf, order_cmpfirst))
                    Js.append(dof_index(blk.cols(), j, M, Kf, order_cmpfirst))
                    Vs.append(blk.nzval)
        return coo_to_csc(np.concatenate(Is), np.concatenate(Js), np.concatenate(Vs), (Kf * M, Kf * M))

    Ce = curl([b == EE for b in boundft], smi, cm, ce)
    Cm = curl([b == HH for b in boundft], sei, ce, cm)
    Pe = paramop(eps, ce, [b != EE for b in boundft], sdl_m, sei)
    Pm = paramop(mu, cm, [b != HH for b in boundft], sdl_e, smi)
    return compose(ft, omega, Pe, Pm, Ce, Cm)
