"""Shapes of the material pipeline (the subset of GeometryPrimitives the GPU kernel covers) and the host side of
`add_obj!` / `clear_objs!` / `calc_matparams!` (reference src/model/model.jl:93-120, src/model/full.jl:16-70).

    add_obj(mdl, "SiO2", Box(c, r), eps=2.085)          # background first: every voxel corner must be covered
    add_obj(mdl, "Si", Box(...), Cylinder(...), eps=12.085)
    calc_matparams(mdl)                                  # fills mdl.eps_arr / mdl.mu_arr on the GPU

`create_paramops(mdl)` of the reference calls calc_matparams! itself (model.jl:143); here it does so when objects
have been added (models whose arrays were filled directly keep them).  There is no CPU path: the arrays come from
`fdfd_calc_matparams` (csrc/matparams.cu)."""
import ctypes as C

import numpy as np

from . import _lib as L
from .grid import EE, HH


class Shape:
    kind = None

    def _fill(self, s):
        raise NotImplementedError


class Box(Shape):
    """Axis-aligned cuboid: centre c, half-widths r (GeometryPrimitives Cuboid with the identity axes)."""
    kind = L.SHAPE_BOX

    def __init__(self, c, r):
        self.c = tuple(float(v) for v in np.atleast_1d(c))
        self.r = tuple(float(v) for v in np.atleast_1d(r))
        if not 1 <= len(self.c) <= 3 or len(self.r) != len(self.c) or min(self.r) < 0:
            raise ValueError("Box(c, r): K centre coordinates and K non-negative half-widths (K = 3; 2 / 1 on 2-D / 1-D models)")
        self.axis = 0


class Ball(Shape):
    """Sphere: centre c, radius."""
    kind = L.SHAPE_BALL

    def __init__(self, c, radius):
        self.c = tuple(float(v) for v in np.atleast_1d(c))
        if not 1 <= len(self.c) <= 3 or radius < 0:
            raise ValueError("Ball(c, radius)")
        self.r = (float(radius), 0.0, 0.0)
        self.axis = 0


Sphere = Ball


class Cylinder(Shape):
    """Right circular cylinder along coordinate axis `axis`: centre c, radius, half-height h."""
    kind = L.SHAPE_CYLINDER

    def __init__(self, c, radius, h, axis=2):
        self.c = tuple(float(v) for v in c)
        if len(self.c) != 3 or radius < 0 or h < 0 or axis not in (0, 1, 2):
            raise ValueError("Cylinder(c, radius, h, axis)")
        self.r = (float(radius), float(h), 0.0)
        self.axis = int(axis)


def extrude(shp, cmp_s, big):
    """The 3-D shape a K-dimensional shape of a 2-D / 1-D model stands for: invariant along the missing axes (half-width
    `big` there, centred on the unit cell [0, 1] of the embedding grid, reduced.py).  A disc becomes a cylinder along
    the missing axis, an interval a slab."""
    K = len(cmp_s)
    if len(shp.c) != K:
        raise ValueError(f"a {K}-D model takes {K}-D shapes")
    if isinstance(shp, Cylinder):
        raise ValueError("Cylinder is a 3-D shape")
    c = [0.5, 0.5, 0.5]
    for k, a in enumerate(cmp_s):
        c[a] = shp.c[k]
    if isinstance(shp, Ball) and K == 2:
        axis = ({0, 1, 2} - set(cmp_s)).pop()
        return Cylinder(c, shp.r[0], big, axis)
    r = [big, big, big]
    for k, a in enumerate(cmp_s):
        r[a] = shp.r[k] if isinstance(shp, Box) else shp.r[0]
    return Box(c, r)


def _as_tensor(p):
    """MatParam of the reference: a scalar, 3 diagonal entries, or a 3x3 tensor"""
    p = np.asarray(p, dtype=np.complex128)
    if p.ndim == 0:
        return np.eye(3, dtype=np.complex128) * p
    if p.shape == (3,):
        return np.diag(p)
    if p.shape == (3, 3):
        return p.copy()
    raise ValueError("material parameter must be a scalar, 3 diagonal entries or a 3x3 tensor")


def clear_objs(mdl):
    """clear_objs! (model.jl:93-104)"""
    mdl.eps_arr[...] = 0
    mdl.mu_arr[...] = 0
    mdl.oind2shp, mdl.oind2epsind, mdl.oind2muind = [], [], []
    mdl.epsind2eps, mdl.muind2mu = [], []


def _param_index(table, P):
    for i, Q in enumerate(table):
        if np.array_equal(P, Q):
            return i
    table.append(P)
    return len(table) - 1


def add_obj(mdl, matname, *shapes, eps=1.0, mu=1.0):
    """add_obj!(mdl, matname, shapes...; ε, μ) (model.jl:107-120): later objects lie on top of earlier ones; equal
    material tensors share one parameter index (a voxel between two objects of the same material is not smoothed)."""
    if len(shapes) == 1 and isinstance(shapes[0], (list, tuple)):
        shapes = tuple(shapes[0])
    if not hasattr(mdl, "oind2shp"):
        mdl.oind2shp, mdl.oind2epsind, mdl.oind2muind, mdl.epsind2eps, mdl.muind2mu = [], [], [], [], []
    Pe, Pm = _as_tensor(eps), _as_tensor(mu)
    for shp in shapes:
        if not isinstance(shp, Shape):
            raise TypeError("add_obj: shapes must be Box / Ball / Cylinder")
        mdl.oind2shp.append(shp)
        mdl.oind2epsind.append(_param_index(mdl.epsind2eps, Pe))
        mdl.oind2muind.append(_param_index(mdl.muind2mu, Pm))


def calc_matparams_array(grid, boundft, ft, shapes, pinds, params, k0=0, k1=None, device=-1, field_ortho_shape=False,
                         julia_layout=False):
    """Smoothed parameter array of field type ft over planes [k0,k1): indexed [i,j,k,v,u] (or, with julia_layout, the
    C-contiguous (3,3,nzl,Ny,Nx) buffer that is the memory of the Julia array and that set_eps takes without a copy)."""
    N = tuple(int(n) for n in grid.N)
    k1 = N[2] if k1 is None else int(k1)
    nzl = k1 - int(k0)
    sh = (L.Shape * len(shapes))()
    for s, shp, pi in zip(sh, shapes, pinds):
        s.kind, s.axis, s.pind = shp.kind, shp.axis, int(pi)
        s.c[:] = shp.c
        s.r[:] = shp.r
    prm = np.ascontiguousarray(np.stack([_as_tensor(P) for P in params]).reshape(-1, 9))
    lprim = [np.ascontiguousarray(a, dtype=np.float64) for a in grid.lg_prim]
    d = L.MatParamsDesc()
    d.N[:] = N
    d.isbloch[:] = [1 if b else 0 for b in grid.isbloch]
    d.boundft_is_E[:] = [1 if b == EE else 0 for b in boundft]
    d.field_type = L.FT_EE if ft == EE else L.FT_HH
    d.field_ortho_shape = 1 if field_ortho_shape else 0
    d.lprim[:] = [a.ctypes.data for a in lprim]
    d.k0, d.k1 = int(k0), k1
    d.nshape, d.nparam = len(shapes), prm.shape[0]
    d.shapes = C.cast(sh, C.c_void_p).value
    d.params = prm.ctypes.data
    d.device = int(device)
    out = np.empty((3, 3, nzl, N[1], N[0]), dtype=np.complex128)
    L.check(L.lib().fdfd_calc_matparams(C.byref(d), out.ctypes.data, L.HOST))
    return out if julia_layout else np.ascontiguousarray(out.transpose(4, 3, 2, 1, 0))


def _calc_matparams_reduced(mdl, device):
    """calc_matparams!(mdl::ModelTE / ModelTM / ModelTEM) (te.jl:17-64, tm.jl:17-64, tem.jl:16-59) on the 3-D kernel:
    the K-dimensional shapes are extruded along the missing axes, the grid is one periodic unit cell thick there, each
    material tensor is reduced to the block of the field's components first (sub_pind2matprm, te.jl:43-44) and
    embedded with 1 on the rest of the diagonal, and the fields that are orthogonal to the shape dimensions (E_z of
    TM, H_z of TE, both fields of TEM) are averaged arithmetically (ise˔shp / ish˔shp, model.jl:65-69).  The block of
    the result is the K-dimensional array: an interface that is invariant along an axis has its normal in the shape
    dimensions, so the 3-D smoothing of the embedded tensor does not mix the block with the rest."""
    from .grid import Grid
    g, cs = mdl.grid, mdl.cmp_s
    K = len(cs)
    big = 1e6 * max(1.0, max(g.L))
    lprim = [np.array([0.0, 1.0])] * 3
    isbloch, boundft = [True] * 3, [EE] * 3
    for k, a in enumerate(cs):
        lprim[a], isbloch[a], boundft[a] = g.lg_prim[k], g.isbloch[k], mdl.boundft[k]
    g3 = Grid(lprim, isbloch)
    shapes3 = [extrude(s, cs, big) for s in mdl.oind2shp]
    for arr, ft, cmps, pinds, table in ((mdl.eps_arr, EE, mdl.cmp_e, mdl.oind2epsind, mdl.epsind2eps),
                                        (mdl.mu_arr, HH, mdl.cmp_m, mdl.oind2muind, mdl.muind2mu)):
        Kf = len(cmps)
        params3 = []
        for P in table:
            Q = np.eye(3, dtype=np.complex128)
            for i, ci in enumerate(cmps):
                for j, cj in enumerate(cmps):
                    Q[ci, cj] = P[ci, cj]
            params3.append(Q)
        if len(table) == 1:                                   # one material: nothing to rasterise
            for i, ci in enumerate(cmps):
                for j, cj in enumerate(cmps):
                    arr[..., i, j] = params3[0][ci, cj]
            continue
        a3 = calc_matparams_array(g3, boundft, ft, shapes3, pinds, params3, device=device, field_ortho_shape=(Kf + K == 3))
        for i, ci in enumerate(cmps):
            for j, cj in enumerate(cmps):
                arr[..., i, j] = a3[..., ci, cj].reshape(g.N)


def calc_matparams(mdl, device=-1):
    """calc_matparams!(mdl) (full.jl:16-70): assignment + subpixel smoothing of eps and mu from the added objects."""
    if not getattr(mdl, "oind2shp", None):
        raise ValueError("calc_matparams: no objects (add_obj) in the model")
    if len(mdl.grid) < 3:
        return _calc_matparams_reduced(mdl, device)
    mdl.eps_arr[...] = calc_matparams_array(mdl.grid, mdl.boundft, EE, mdl.oind2shp, mdl.oind2epsind, mdl.epsind2eps,
                                            device=device)
    if len(mdl.muind2mu) == 1 and np.array_equal(mdl.muind2mu[0], np.eye(3)):
        mdl.mu_arr[...] = 0
        for v in range(3):
            mdl.mu_arr[..., v, v] = 1.0       # one material with mu = 1: nothing to rasterise
    else:
        mdl.mu_arr[...] = calc_matparams_array(mdl.grid, mdl.boundft, HH, mdl.oind2shp, mdl.oind2muind, mdl.muind2mu,
                                               device=device)
