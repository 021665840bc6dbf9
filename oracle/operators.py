"""Sparse-matrix restatement of the reference's operator assembly (TEST INFRASTRUCTURE ONLY).

Reference path restated here:
  create_curls      src/model/model.jl:160-175   -> create_curl   (StaggeredGridCalculus, not in tree)
  create_paramops   src/model/model.jl:141-158   -> create_paramop / create_mean (MaxwellBase, not in tree)
  create_A          src/model/model.jl:225-246      A = Cm*(Pmu \\ Ce) - w^2*Peps   (EE)
                                                    A = Ce*(Peps \\ Cm) - w^2*Pmu   (HH)
The out-of-tree pieces follow SURVEY.md Appendix A.2-A.6 ("parity unpinned", see oracle/__init__).

Every sparse matrix is carried as a `Csc` (colptr,rowval,nzval) triple built by our OWN
COO->CSC routine that mimics Julia's `sparse(I,J,V)`: duplicates summed, rows sorted within a
column, explicit zeros KEPT (scipy drops them inconsistently).  Index arrays are
0-based internally; `julia_pattern()` returns the 1-based Int64 pair the debug export must match.
"""
from __future__ import annotations

from dataclasses import dataclass
import numpy as np
import scipy.sparse as sp

from .grid import EE, HH

# --------------------------------------------------------------------------------------
# CSC container that keeps explicit zeros
# --------------------------------------------------------------------------------------


@dataclass
class Csc:
    shape: tuple
    colptr: np.ndarray  # int64, len ncol+1, 0-based
    rowval: np.ndarray  # int64, len nnz, 0-based, sorted within each column
    nzval: np.ndarray   # complex128, len nnz (may hold explicit zeros)

    @property
    def nnz(self):
        return int(self.rowval.size)

    def to_scipy(self):
        """scipy view for numerics (stored zeros are harmless for products with vectors)."""
        return sp.csc_matrix((self.nzval, self.rowval, self.colptr), shape=self.shape)

    def julia_pattern(self):
        """(colptr, rowval) as Julia stores them: 1-based Int64."""
        return self.colptr.astype(np.int64) + 1, self.rowval.astype(np.int64) + 1

    def cols(self):
        return np.repeat(np.arange(self.shape[1], dtype=np.int64), np.diff(self.colptr))

    def matvec(self, x):
        return self.to_scipy() @ x


def coo_to_csc(I, J, V, shape) -> Csc:
    """Julia `sparse(I,J,V,m,n)`: sum duplicates, sort rows within columns, keep explicit zeros."""
    I = np.asarray(I, dtype=np.int64).ravel()
    J = np.asarray(J, dtype=np.int64).ravel()
    V = np.asarray(V, dtype=np.complex128).ravel()
    m, n = shape
    if I.size == 0:
        return Csc(shape, np.zeros(n + 1, np.int64), np.zeros(0, np.int64), np.zeros(0, np.complex128))
    key = J * m + I
    order = np.argsort(key, kind="stable")
    key, V = key[order], V[order]
    first = np.concatenate(([True], key[1:] != key[:-1]))
    starts = np.flatnonzero(first)
    vals = np.add.reduceat(V, starts)
    ukey = key[starts]
    rows, cols = ukey % m, ukey // m
    colptr = np.zeros(n + 1, np.int64)
    np.add.at(colptr, cols + 1, 1)
    colptr = np.cumsum(colptr)
    return Csc(shape, colptr, rows, vals)


def _struct_ones(A: Csc):
    """scipy matrix with value 1 on every STORED entry (explicit zeros included)."""
    return sp.csc_matrix((np.ones(A.nnz), A.rowval, A.colptr), shape=A.shape)


def spgemm(A: Csc, B: Csc) -> Csc:
    """Julia `A*B` for SparseMatrixCSC: the result holds the STRUCTURAL product pattern
    (entries whose numerical sum is 0 stay stored) - SURVEY A.6 rule 3."""
    S = (_struct_ones(A) @ _struct_ones(B)).tocoo()          # positive sums never cancel
    Vm = (A.to_scipy() @ B.to_scipy()).tocoo()               # may have dropped zeros
    I = np.concatenate((S.row, Vm.row))
    J = np.concatenate((S.col, Vm.col))
    V = np.concatenate((np.zeros(S.nnz, np.complex128), Vm.data))
    return coo_to_csc(I, J, V, (A.shape[0], B.shape[1]))


def diag_ldiv(P: Csc, B: Csc) -> Csc:
    """Julia `P \\ B` for a (numerically) diagonal sparse P: rows of B rescaled, pattern of B
    unchanged (SURVEY A.6 rule 2).  Raises like the reference would for a non-diagonal P
    (model.jl:236 comment: division unsupported when P is not diagonal)."""
    Ps = P.to_scipy().tocoo()
    off = (Ps.row != Ps.col) & (Ps.data != 0)
    if off.any():
        raise ValueError("P \\ B: P must be diagonal (reference model.jl:236)")
    d = np.zeros(P.shape[0], np.complex128)
    on = Ps.row == Ps.col
    np.add.at(d, Ps.row[on], Ps.data[on])
    return Csc(B.shape, B.colptr.copy(), B.rowval.copy(), B.nzval / d[B.rowval])


def sub_scaled(A: Csc, alpha, B: Csc) -> Csc:
    """Julia `A - alpha*B` for sparse A, B: union of stored entries, results that are exactly
    zero are DROPPED (map-based subtraction; SURVEY A.6 rule 4)."""
    I = np.concatenate((A.rowval, B.rowval))
    J = np.concatenate((A.cols(), B.cols()))
    V = np.concatenate((A.nzval, -(alpha * B.nzval)))
    C = coo_to_csc(I, J, V, A.shape)
    keep = C.nzval != 0
    cols = C.cols()[keep]
    colptr = np.zeros(C.shape[1] + 1, np.int64)
    np.add.at(colptr, cols + 1, 1)
    return Csc(C.shape, np.cumsum(colptr), C.rowval[keep], C.nzval[keep])


# --------------------------------------------------------------------------------------
# DOF numbering (reference model.jl:75-83)
# --------------------------------------------------------------------------------------


def dof_index(cell, cmp, ncell, ncmp, order_cmpfirst=True):
    """0-based DOF index of component `cmp` at linear cell index `cell` (x fastest)."""
    return ncmp * cell + cmp if order_cmpfirst else ncell * cmp + cell


def field_arr2vec(F, order_cmpfirst=True):
    """MaxwellBase.field_arr2vec (call site model.jl:203-204): F[i,j,k,c] -> DOF vector."""
    F = np.asarray(F)
    if order_cmpfirst:
        return np.ascontiguousarray(F.transpose(2, 1, 0, 3)).ravel()
    return np.ascontiguousarray(F.transpose(3, 2, 1, 0)).ravel()


def field_vec2arr(v, N, ncmp=3, order_cmpfirst=True):
    Nx, Ny, Nz = N
    if order_cmpfirst:
        return np.asarray(v).reshape(Nz, Ny, Nx, ncmp).transpose(2, 1, 0, 3)
    return np.asarray(v).reshape(ncmp, Nz, Ny, Nx).transpose(3, 2, 1, 0)


# --------------------------------------------------------------------------------------
# create_d (create_∂) and create_mean  
# --------------------------------------------------------------------------------------


def _cell_index_array(N):
    M = int(np.prod(N))
    return np.arange(M, dtype=np.int64).reshape(tuple(N), order="F")  # x fastest


def _slab(N, nw, i):
    sl = [slice(None)] * len(N)
    sl[nw] = i
    return tuple(sl)


def create_d_info(nw, isfwd, N, dw_inv, isbloch, e_mikL):
    """COO triplets of create_∂ on the scalar cell grid.
    Rows/cols are linear cell indices.  Two stored entries per row: the 'diagonal' one
    (same cell) and the 'shifted' one (cell +1 for forward, -1 for backward, periodic wrap);
    symmetry boundaries overwrite some VALUES with explicit zeros, never the positions."""
    N = tuple(int(n) for n in N)
    Nw = N[nw]
    idx = _cell_index_array(N)
    Js = np.roll(idx, -1 if isfwd else +1, axis=nw)  # Js[i] = idx[i+1] (fwd) / idx[i-1] (bwd)
    shp = [1] * len(N)
    shp[nw] = Nw
    DW = np.asarray(dw_inv, dtype=np.complex128).reshape(shp)
    sgn = 1.0 if isfwd else -1.0
    V0 = (-sgn) * np.ones(N, np.complex128) * DW
    Vs = (+sgn) * np.ones(N, np.complex128) * DW
    if isbloch:
        if isfwd:
            Vs[_slab(N, nw, Nw - 1)] *= e_mikL      # f[N+1] := e^{-ikL} f[1]
        else:
            Vs[_slab(N, nw, 0)] /= e_mikL           # g[0] := g[N] / e^{-ikL}
    else:
        if isfwd:
            V0[_slab(N, nw, 0)] = 0                 # field on the symmetry boundary vanishes
            Vs[_slab(N, nw, Nw - 1)] = 0            # ghost f[N+1] := 0
        else:
            V0[_slab(N, nw, 0)] = 0                 # even image => zero derivative in row 1
            Vs[_slab(N, nw, 0)] = 0
    I = np.concatenate((idx.ravel(order="F"), idx.ravel(order="F")))
    J = np.concatenate((idx.ravel(order="F"), Js.ravel(order="F")))
    V = np.concatenate((V0.ravel(order="F"), Vs.ravel(order="F")))
    return I, J, V


def create_d(nw, isfwd, N, dw_inv=None, isbloch=True, e_mikL=1.0) -> Csc:
    N = tuple(int(n) for n in N)
    if dw_inv is None:
        dw_inv = np.ones(N[nw])
    M = int(np.prod(N))
    return coo_to_csc(*create_d_info(nw, isfwd, N, dw_inv, isbloch, e_mikL), (M, M))


def create_m_info(nw, isfwd, N, dw=None, dw_out_inv=None, isbloch=True, e_mikL=1.0):
    """COO triplets of create_mean along axis nw on the scalar cell grid.
    dw: cell sizes at the INPUT locations, dw_out_inv: reciprocal sizes at the OUTPUT locations;
    both None => unweighted arithmetic mean."""
    N = tuple(int(n) for n in N)
    Nw = N[nw]
    idx = _cell_index_array(N)
    shift = -1 if isfwd else +1
    Js = np.roll(idx, shift, axis=nw)
    dw = np.ones(Nw, np.complex128) if dw is None else np.asarray(dw, np.complex128)
    dwo = np.ones(Nw, np.complex128) if dw_out_inv is None else np.asarray(dw_out_inv, np.complex128)
    shp = [1] * len(N)
    shp[nw] = Nw
    w0 = (0.5 * dwo * dw).reshape(shp)
    ws = (0.5 * dwo * np.roll(dw, shift)).reshape(shp)
    V0 = np.ones(N, np.complex128) * w0
    Vs = np.ones(N, np.complex128) * ws
    if isbloch:
        if isfwd:
            Vs[_slab(N, nw, Nw - 1)] *= e_mikL
        else:
            Vs[_slab(N, nw, 0)] /= e_mikL
    else:
        if isfwd:
            V0[_slab(N, nw, 0)] = 0
            Vs[_slab(N, nw, Nw - 1)] = 0
        else:
            V0[_slab(N, nw, 0)] *= 2                # g[0] := g[1] (even image): "0's and 2's"
            Vs[_slab(N, nw, 0)] = 0
    I = np.concatenate((idx.ravel(order="F"), idx.ravel(order="F")))
    J = np.concatenate((idx.ravel(order="F"), Js.ravel(order="F")))
    V = np.concatenate((V0.ravel(order="F"), Vs.ravel(order="F")))
    return I, J, V


def create_m(nw, isfwd, N, dw=None, dw_out_inv=None, isbloch=True, e_mikL=1.0) -> Csc:
    N = tuple(int(n) for n in N)
    M = int(np.prod(N))
    return coo_to_csc(*create_m_info(nw, isfwd, N, dw, dw_out_inv, isbloch, e_mikL), (M, M))


# --------------------------------------------------------------------------------------
# create_curl, create_mean (3-component, used by create_Mcs), create_paramop (A.5)
# --------------------------------------------------------------------------------------


def _levi_civita(v, w, u):
    return (v - w) * (w - u) * (u - v) / 2  # +1 / -1 / 0 for 0-based distinct indices


def create_curl(isfwd, dl_inv, isbloch, e_mikL, order_cmpfirst=True) -> Csc:
    """3-D curl: block (v,u) = eps_{v w u} * create_d(w, isfwd[w], ...), w = 3-u-v (0-based)."""
    N = tuple(len(a) for a in dl_inv)
    M = int(np.prod(N))
    Is, Js, Vs = [], [], []
    for v in range(3):
        for u in range(3):
            if u == v:
                continue
            w = 3 - u - v
            s = _levi_civita(v, w, u)
            I, J, V = create_d_info(w, bool(isfwd[w]), N, dl_inv[w], bool(isbloch[w]), e_mikL[w])
            Is.append(dof_index(I, v, M, 3, order_cmpfirst))
            Js.append(dof_index(J, u, M, 3, order_cmpfirst))
            Vs.append(s * V)
    return coo_to_csc(np.concatenate(Is), np.concatenate(Js), np.concatenate(Vs), (3 * M, 3 * M))


def create_mean(isfwd, dl, dl_out_inv, isbloch, e_mikL, order_cmpfirst=True) -> Csc:
    """3-component block-diagonal averaging operator (call site model.jl:302-303, create_Mcs):
    component w is averaged along its own axis w."""
    N = tuple(len(a) for a in dl)
    M = int(np.prod(N))
    Is, Js, Vs = [], [], []
    for w in range(3):
        I, J, V = create_m_info(w, bool(isfwd[w]), N, dl[w], dl_out_inv[w], bool(isbloch[w]), e_mikL[w])
        Is.append(dof_index(I, w, M, 3, order_cmpfirst))
        Js.append(dof_index(J, w, M, 3, order_cmpfirst))
        Vs.append(V)
    return coo_to_csc(np.concatenate(Is), np.concatenate(Js), np.concatenate(Vs), (3 * M, 3 * M))


def _place_block(B: Csc, v, u, M, order_cmpfirst):
    return (dof_index(B.rowval, v, M, 3, order_cmpfirst),
            dof_index(B.cols(), u, M, 3, order_cmpfirst), B.nzval)


def create_paramop(param, isfwd_in=None, dl=None, dl_out_inv=None, isbloch=None, e_mikL=None,
                   order_cmpfirst=True, weighted_out=False, diag_only=None) -> Csc:
    """Material operator.  param[i,j,k,v,u]: diagonal entries at the field-component
    locations, off-diagonal entries at the voxel corners (evidence full.jl:58-59,64-65).
        (P F)_v = p_vv .* F_v + sum_{u != v} Mout_v [ p_vu .* ( Min_u F_u ) ]
    Min_u  = create_mean along u, isfwd = isfwd_in[u], weights (dl, dl_out_inv)  (model.jl:149,153)
    Mout_v = create_mean along v, isfwd = !isfwd_in[v]; unweighted unless weighted_out
             (the weighting of the OUTPUT average is the least certain part of the restatement,
             SURVEY A.3, and is kept pluggable: weighted_out=True uses the transposed weights
             (dl_out at the corner as input size, 1/dl at the field point as output size)).
    diag_only=None: auto (True iff every off-diagonal entry of param is exactly zero)."""
    param = np.asarray(param, dtype=np.complex128)
    N = param.shape[:3]
    M = int(np.prod(N))
    cell = np.arange(M, dtype=np.int64)
    Is, Js, Vs = [], [], []
    for v in range(3):
        Is.append(dof_index(cell, v, M, 3, order_cmpfirst))
        Js.append(dof_index(cell, v, M, 3, order_cmpfirst))
        Vs.append(param[..., v, v].ravel(order="F"))
    if diag_only is None:
        off = param.copy()
        for v in range(3):
            off[..., v, v] = 0
        diag_only = not off.any()
    if not diag_only:
        for v in range(3):
            for u in range(3):
                if u == v:
                    continue
                Min = create_m(u, bool(isfwd_in[u]), N, dl[u], dl_out_inv[u], bool(isbloch[u]), e_mikL[u])
                if weighted_out:
                    Mout = create_m(v, not bool(isfwd_in[v]), N, 1.0 / np.asarray(dl_out_inv[v]),
                                    1.0 / np.asarray(dl[v]), bool(isbloch[v]), e_mikL[v])
                else:
                    Mout = create_m(v, not bool(isfwd_in[v]), N, None, None, bool(isbloch[v]), e_mikL[v])
                pvu = param[..., v, u].ravel(order="F")
                # diag(pvu) * Min : scale the rows of Min
                scaled = Csc(Min.shape, Min.colptr, Min.rowval, Min.nzval * pvu[Min.rowval])
                blk = spgemm(Mout, scaled)
                I, J, V = _place_block(blk, v, u, M, order_cmpfirst)
                Is.append(I), Js.append(J), Vs.append(V)
    return coo_to_csc(np.concatenate(Is), np.concatenate(Js), np.concatenate(Vs), (3 * M, 3 * M))


# --------------------------------------------------------------------------------------
# Reference-level composition (model.jl:141-175, 225-274)
# --------------------------------------------------------------------------------------


def create_curls(sdl_e_inv, sdl_m_inv, boundft, isbloch, e_mikL, order_cmpfirst=True):
    """Reference create_curls (model.jl:160-175): Ce uses isfwd = boundft.==EE and 1/sdl_m,
    Cm uses isfwd = boundft.==HH and 1/sdl_e."""
    isfwd_e = [b == EE for b in boundft]
    isfwd_m = [b == HH for b in boundft]
    Ce = create_curl(isfwd_e, sdl_m_inv, isbloch, e_mikL, order_cmpfirst)
    Cm = create_curl(isfwd_m, sdl_e_inv, isbloch, e_mikL, order_cmpfirst)
    return Ce, Cm


def create_paramops(eps, mu, sdl_e, sdl_m, sdl_e_inv, sdl_m_inv, boundft, isbloch, e_mikL,
                    order_cmpfirst=True, weighted_out=False):
    """Reference create_paramops (model.jl:141-158) minus calc_matparams! (eps/mu arrays are inputs)."""
    isfwd_in_e = [b != EE for b in boundft]
    isfwd_in_m = [b != HH for b in boundft]
    Pe = create_paramop(eps, isfwd_in_e, sdl_m, sdl_e_inv, isbloch, e_mikL, order_cmpfirst, weighted_out)
    Pm = create_paramop(mu, isfwd_in_m, sdl_e, sdl_m_inv, isbloch, e_mikL, order_cmpfirst, weighted_out)
    return Pe, Pm


def create_A(ft, omega, Pe: Csc, Pm: Csc, Ce: Csc, Cm: Csc) -> Csc:
    """Reference create_A (model.jl:225-246)."""
    if ft == EE:
        A = spgemm(Cm, diag_ldiv(Pm, Ce))
        if omega != 0:
            A = sub_scaled(A, omega ** 2, Pe)
    elif ft == HH:
        A = spgemm(Ce, diag_ldiv(Pe, Cm))
        if omega != 0:
            A = sub_scaled(A, omega ** 2, Pm)
    else:
        raise ValueError(f"ft = {ft} is unsupported.")
    return A


def create_b(ft, omega, Pe: Csc, Pm: Csc, Ce: Csc, Cm: Csc, je, jm):
    """Reference create_b (model.jl:251-274): EE: b = -Cm (Pmu \\ jm) - i w je."""
    je = np.asarray(je, np.complex128)
    jm = np.asarray(jm, np.complex128)
    if ft == EE:
        d = Pm.to_scipy().diagonal()
        b = -(Cm.matvec(jm / d))
        if omega != 0:
            b = b - (1j * omega) * je
    elif ft == HH:
        d = Pe.to_scipy().diagonal()
        b = Ce.matvec(je / d)
        if omega != 0:
            b = b - (1j * omega) * jm
    else:
        raise ValueError(f"ft = {ft} is unsupported.")
    return b


def h_from_e(e, omega, Pm: Csc, Ce: Csc, jm):
    """Reference h_from_e (model.jl:276-279): (i/w) * (Pmu \\ (Ce e + jm))."""
    d = Pm.to_scipy().diagonal()
    return (1j / omega) * ((Ce.matvec(e) + jm) / d)


def e_from_h(h, omega, Pe: Csc, Cm: Csc, je):
    """Reference e_from_h (model.jl:281-284): (-i/w) * (Peps \\ (Cm h - je)); diagonal Peps only."""
    d = Pe.to_scipy().diagonal()
    return (-1j / omega) * ((Cm.matvec(h) - je) / d)


def create_Mcs(sdl_e, sdl_m, sdl_e_inv, sdl_m_inv, boundft, isbloch, e_mikL, order_cmpfirst=True):
    """Reference create_Mcs (model.jl:287-306): operators interpolating fields to voxel corners."""
    isfwd_e = [b != EE for b in boundft]
    isfwd_m = [b != HH for b in boundft]
    Mce = create_mean(isfwd_e, sdl_m, sdl_e_inv, isbloch, e_mikL, order_cmpfirst)
    Mcm = create_mean(isfwd_m, sdl_e, sdl_m_inv, isbloch, e_mikL, order_cmpfirst)
    return Mce, Mcm
