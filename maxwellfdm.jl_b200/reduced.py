"""2-D and 1-D models of the reference (ModelTE te.jl:4-14, ModelTM tm.jl:4-14, ModelTEM tem.jl:4-13) on the 3-D
GPU path.

The reference builds these from the same create_curl / create_paramop with `cmp_shp`, `cmp_out`, `cmp_in`
(model.jl:171-172): the grid has K < 3 axes (`cmpₛ`), E has the components `cmpₑ`, H the components `cmpₘ`.
A K-dimensional problem is the 3-D problem that is one cell thick and periodic (phase 1) along the missing axes: with
∆ = 1 there the difference stencil is exactly (+1)·f + (−1)·f = 0, the 3-D operator decouples into the TE and TM
blocks, and the block of the model's components IS the reference's K-dimensional operator.  So the reduced model
runs on the same kernels (libfdfd_b200, 3-D handle with N = 1 along the missing axes); this module only embeds the
K-dimensional vectors into the 3-component layout and extracts the result.  Krylov iterates started in one block stay
in it exactly (the other block's entries are exact zeros), so `solve` is the K-dimensional solve.
Objects: shapes.calc_matparams extrudes the K-dimensional shapes the same way (material kernel on the 3-D scene).
Not covered: z-slabs.
"""
import numpy as np

from .grid import EE, HH


def embed_geometry(g):
    """3-D (N, isbloch, sdl_e, sdl_m, e_mikL, boundft) of a K-dimensional description (model._Geom)."""
    N, isbloch, sdl_e, sdl_m, ph, bft = [], [], [], [], [], []
    for a in range(3):
        if a in g.cmp_s:
            k = g.cmp_s.index(a)
            N.append(g.N[k]); isbloch.append(g.isbloch[k]); sdl_e.append(g.sdl_e[k]); sdl_m.append(g.sdl_m[k])
            ph.append(g.e_mikL[k]); bft.append(g.boundft[k])
        else:   # missing axis: one cell, periodic with phase 1, ∆ = 1 (exact cancellation of its differences)
            N.append(1); isbloch.append(True); sdl_e.append(np.ones(1, complex)); sdl_m.append(np.ones(1, complex))
            ph.append(1.0 + 0j); bft.append(EE)
    return tuple(N), tuple(isbloch), tuple(sdl_e), tuple(sdl_m), np.asarray(ph, complex), tuple(bft)


def embed_param(arr, N3, cmps):
    """(N..., Kf, Kf) material array -> (Nx,Ny,Nz,3,3): the model's block, 1 on the rest of the diagonal.  An array
    that was never assigned (all zeros) or is the identity gives None (parameter == 1)."""
    Kf = len(cmps)
    arr = np.asarray(arr)
    ident = np.zeros((Kf, Kf))
    np.fill_diagonal(ident, 1.0)
    if not arr.any() or np.array_equal(arr, np.broadcast_to(ident, arr.shape)):
        return None
    out = np.zeros(tuple(N3) + (3, 3), np.complex128)
    for v in range(3):
        out[..., v, v] = 1.0
    for i, ci in enumerate(cmps):
        for j, cj in enumerate(cmps):
            out[..., ci, cj] = arr[..., i, j].reshape(N3)
    return out


class ReducedOperator:
    """The K-dimensional operator of ModelTE / ModelTM / ModelTEM: same methods as FdfdOperator on vectors of
    Kf * prod(N) entries (reference DOF order, model.jl:75-83 with Kf components)."""

    def __init__(self, A3, ncell, cmp_e, cmp_m, ft, order_cmpfirst=True):
        self.A3, self.ncell, self.ft, self.order_cmpfirst = A3, int(ncell), ft, bool(order_cmpfirst)
        self.cmp_e, self.cmp_m = tuple(cmp_e), tuple(cmp_m)
        self.cmp_f = self.cmp_e if ft == EE else self.cmp_m          # components of the unknown
        self.n = self.ncell * len(self.cmp_f)

    # -- layout -----------------------------------------------------------------------------------
    def _embed(self, v, cmps, name="x"):
        if v is None:
            return None
        K, nc = len(cmps), self.ncell
        if hasattr(v, "is_cuda"):
            import torch
            if v.dtype != torch.complex128 or v.numel() != K * nc:
                raise ValueError(f"{name} must be a complex128 tensor of {K * nc} elements")
            x3 = torch.zeros(3 * nc, dtype=torch.complex128, device=v.device)
            src = v.reshape(nc, K) if self.order_cmpfirst else v.reshape(K, nc).t()
            x3.view(nc, 3)[:, list(cmps)] = src
            return x3
        v = np.asarray(v, dtype=np.complex128)
        if v.shape != (K * nc,):
            raise ValueError(f"{name} must have {K * nc} elements")
        x3 = np.zeros((nc, 3), np.complex128)
        x3[:, list(cmps)] = v.reshape(nc, K) if self.order_cmpfirst else v.reshape(K, nc).T
        return x3.reshape(-1)

    def _extract(self, v3, cmps):
        K, nc = len(cmps), self.ncell
        if hasattr(v3, "is_cuda"):
            sub = v3.view(nc, 3)[:, list(cmps)]
            return (sub if self.order_cmpfirst else sub.t()).contiguous().reshape(-1)
        sub = np.asarray(v3).reshape(nc, 3)[:, list(cmps)]
        return np.ascontiguousarray(sub if self.order_cmpfirst else sub.T).reshape(-1)

    # -- operator ---------------------------------------------------------------------------------
    def __matmul__(self, x):
        return self._extract(self.A3 @ self._embed(x, self.cmp_f), self.cmp_f)

    __mul__ = __matmul__

    def mul(self, y, x, transpose=False):
        """mul!(y, A, x)"""
        x3 = self._embed(x, self.cmp_f)
        r = self._extract(self.A3.rmatvec_T(x3) if transpose else self.A3 @ x3, self.cmp_f)
        if hasattr(y, "copy_"):
            y.copy_(r)
        else:
            y[...] = r
        return y

    def rmatvec_T(self, x):
        return self._extract(self.A3.rmatvec_T(self._embed(x, self.cmp_f)), self.cmp_f)

    def solve(self, b, x0=None, **kw):
        """x = A \\ b (BiCGSTAB / QMR on the device); returns (x, info)."""
        x3, info = self.A3.solve(self._embed(b, self.cmp_f, "b"), self._embed(x0, self.cmp_f, "x0"), **kw)
        return self._extract(x3, self.cmp_f), info

    def create_b(self, je, jm=None):
        """EE: b = -Cm(Pmu \\ jm) - iω je ; HH: b = Ce(Peps \\ je) - iω jm   (model.jl:251-274)"""
        b3 = self.A3.create_b(self._embed(je, self.cmp_e, "je"), self._embed(jm, self.cmp_m, "jm"))
        return self._extract(b3, self.cmp_f)

    def h_from_e(self, e, jm=None):
        return self._extract(self.A3.h_from_e(self._embed(e, self.cmp_e, "e"), self._embed(jm, self.cmp_m, "jm")), self.cmp_m)

    def e_from_h(self, h, je=None):
        return self._extract(self.A3.e_from_h(self._embed(h, self.cmp_m, "h"), self._embed(je, self.cmp_e, "je")), self.cmp_e)

    def interp_corners(self, f, ft="E"):
        cm = self.cmp_e if (str(ft).upper().startswith("E") or ft == 0) else self.cmp_m
        return self._extract(self.A3.interp_corners(self._embed(f, cm, "f"), ft), cm)

    def export_pattern(self, values=True):
        """(colptr, rowval, nzval) of the K-dimensional A as Julia stores it (1-based Int64): the block of the 3-D
        export (fdfd_export_pattern) that belongs to the model's components, renumbered to the K-dimensional DOF
        order.  The couplings through the missing axes fall on entries the K-dimensional operator stores anyway
        (they are differences of a cell with itself), so the block has the K-dimensional structure."""
        colptr3, rowval3, nz3 = self.A3.export_pattern(values)
        nc, Kf, cm = self.ncell, len(self.cmp_f), list(self.cmp_f)
        col3 = np.repeat(np.arange(3 * nc, dtype=np.int64), np.diff(colptr3))
        row3 = rowval3 - 1
        slot = np.full(3, -1, np.int64)
        slot[cm] = np.arange(Kf)
        kr, kc = slot[row3 % 3], slot[col3 % 3]
        keep = (kr >= 0) & (kc >= 0)
        cell_r, cell_c, kr, kc = row3[keep] // 3, col3[keep] // 3, kr[keep], kc[keep]
        if self.order_cmpfirst:
            r, c = kr + Kf * cell_r, kc + Kf * cell_c
        else:
            r, c = kr * nc + cell_r, kc * nc + cell_c
        order = np.lexsort((r, c))                      # columns ascending, rows ascending inside a column
        colptr = np.zeros(self.n + 1, np.int64)
        np.add.at(colptr, c + 1, 1)
        return np.cumsum(colptr) + 1, r[order] + 1, (nz3[keep][order] if values else None)

    # -- life cycle / bookkeeping ---------------------------------------------------------------------
    def close(self):
        self.A3.close()

    @property
    def closed(self):
        return self.A3.closed

    @property
    def launch_count(self):
        return self.A3.launch_count
