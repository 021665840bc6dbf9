"""Host-side mirror of the reference's model API for the hot path (reference src/model/model.jl).

Same names and argument meaning as the reference (ASCII spellings: ω→w, ε→eps, μ→mu, ∆→d, ₑ→e, ₘ→m):
    ModelFull(grid)                     full.jl:7-13
    ModelTE / ModelTM / ModelTEM        te.jl:4-14, tm.jl:4-14, tem.jl:4-13   (K < 3 grids: reduced.py)
    set_wpml / set_boundft / set_Npml / set_kbloch      model.jl:86-89
    create_e_mikL                       model.jl:91
    clear_srcs / add_srce / add_srcm / create_srcs      model.jl:177-207
    create_stretched_dls                model.jl:122-139
    create_paramops / create_curls      model.jl:141-175   (descriptions, not sparse matrices)
    create_A / create_b / create_linsys model.jl:209-274   (GPU: returns an FdfdOperator, not a CSC)
    h_from_e / e_from_h / create_Mcs    model.jl:276-306
    solve                               (absent in the reference: `A \\ b` left to the user)
The reference call sequence carries over: `Ps = create_paramops(mdl); Cs = create_curls(mdl); js = create_srcs(mdl);
A, b = create_linsys(EE, ω, Ps, Cs, js)`, then `e, info = solve(A, b)` instead of `e = A \\ b`, then
`h = h_from_e(e, ω, Ps, Cs, js)`.  The operator is never assembled.  Shortcuts taking the model are kept:
`create_linsys(EE, ω, mdl)`.
Geometry rasterisation + subpixel smoothing (calc_matparams!, full.jl:16-70): shapes.py / csrc/matparams.cu; models
without objects may fill mdl.eps_arr / mdl.mu_arr directly.
"""
import itertools

import numpy as np

from .grid import EE, HH, PRIM, DUAL, Grid, PMLParam, create_stretched_dl, ft2gt
from .operator import FdfdOperator, MultiGpuOperator
from .sources import Source
from .reduced import ReducedOperator

_OPS = (FdfdOperator, ReducedOperator, MultiGpuOperator)   # what create_A returns (3-D model / ModelTE, ModelTM, ModelTEM / ngpu=N)


class Model:
    def __init__(self, grid, cmp_s=(0, 1, 2), cmp_e=(0, 1, 2), cmp_m=(0, 1, 2)):
        K = len(grid)
        if len(cmp_s) != K:
            raise ValueError("cmp_s must name one Cartesian axis per grid axis")
        # Cartesian components (0-based) of the shape dimensions, of E and of H (model.jl:41-43)
        self.cmp_s, self.cmp_e, self.cmp_m = tuple(cmp_s), tuple(cmp_e), tuple(cmp_m)
        Ke, Km = len(cmp_e), len(cmp_m)
        self.wpml = 0.0                                   # ωpml (model.jl:37)
        self.grid = grid
        self.boundft = (EE,) * K                          # model.jl:46
        self.Npml = ((0,) * K, (0,) * K)                  # model.jl:47
        self.kbloch = (0.0,) * K                          # model.jl:48
        self.eps_arr = np.zeros(grid.N + (Ke, Ke), np.complex128)   # create_param_array (model.jl:51-52)
        self.mu_arr = np.zeros(grid.N + (Km, Km), np.complex128)
        self.je_arr = np.zeros(grid.N + (Ke,), np.complex128)      # create_field_array (model.jl:55-56)
        self.jm_arr = np.zeros(grid.N + (Km,), np.complex128)
        self.order_cmpfirst = True                        # model.jl:72
        self.pml = PMLParam()

    def size(self, ft):
        Kf = len(self.cmp_e) if ft == EE else len(self.cmp_m)
        return ((Kf,) + self.grid.N) if self.order_cmpfirst else (self.grid.N + (Kf,))   # model.jl:75-81

    def length(self, ft):
        return int(np.prod(self.size(ft)))


def ModelFull(grid):
    if len(grid) != 3:
        raise ValueError("ModelFull needs a 3-D grid")
    return Model(grid)


def ModelTE(grid):
    """2-D TE (te.jl:4-14): shapes in the x-y plane, E = (Ex, Ey), H = (Hz)."""
    if len(grid) != 2:
        raise ValueError("ModelTE needs a 2-D grid")
    return Model(grid, cmp_s=(0, 1), cmp_e=(0, 1), cmp_m=(2,))


def ModelTM(grid):
    """2-D TM (tm.jl:4-14): shapes in the x-y plane, E = (Ez), H = (Hx, Hy)."""
    if len(grid) != 2:
        raise ValueError("ModelTM needs a 2-D grid")
    return Model(grid, cmp_s=(0, 1), cmp_e=(2,), cmp_m=(0, 1))


def ModelTEM(grid):
    """1-D TEM (tem.jl:4-13): shapes along z, E = (Ex), H = (Hy)."""
    if len(grid) != 1:
        raise ValueError("ModelTEM needs a 1-D grid")
    return Model(grid, cmp_s=(2,), cmp_e=(0,), cmp_m=(1,))


def set_wpml(mdl, wpml):
    mdl.wpml = wpml


def set_boundft(mdl, boundft):
    if len(boundft) != len(mdl.grid):
        raise ValueError("boundft must have one entry per axis")
    mdl.boundft = tuple(boundft)


def set_Npml(mdl, Npml):
    mdl.Npml = (tuple(int(n) for n in Npml[0]), tuple(int(n) for n in Npml[1]))


def set_kbloch(mdl, kbloch):
    mdl.kbloch = tuple(float(k) for k in kbloch)


def create_e_mikL(mdl):
    return np.exp(-1j * np.asarray(mdl.kbloch) * np.asarray(mdl.grid.L))


def clear_srcs(mdl):
    mdl.je_arr[...] = 0
    mdl.jm_arr[...] = 0


def _add_src(jarr, ft, mdl, src):
    if not isinstance(src, Source):
        raise TypeError("src must be a Source")
    g = mdl.grid
    src.add(jarr, tuple(ft2gt(ft, b) for b in mdl.boundft), g.bounds, g.l, g.dl, g.isbloch)


def add_srce(mdl, src):
    _add_src(mdl.je_arr, EE, mdl, src)


def add_srcm(mdl, src):
    _add_src(mdl.jm_arr, HH, mdl, src)


def field_arr2vec(F, order_cmpfirst=True):
    """(N..., Kf) array -> DOF vector in the reference's order (model.jl:75-83; first grid axis fastest)."""
    F = np.asarray(F)
    K = F.ndim - 1
    rev = tuple(range(K - 1, -1, -1))
    axes = rev + (K,) if order_cmpfirst else (K,) + rev
    return np.ascontiguousarray(F.transpose(axes)).reshape(-1)


def field_vec2arr(v, N, order_cmpfirst=True):
    N = tuple(int(n) for n in N)
    K = len(N)
    v = np.asarray(v)
    Kf = v.size // int(np.prod(N))
    rev = tuple(range(K - 1, -1, -1))
    if order_cmpfirst:
        return v.reshape(N[::-1] + (Kf,)).transpose(rev + (K,))
    return v.reshape((Kf,) + N[::-1]).transpose(tuple(range(K, 0, -1)) + (0,))


def create_srcs(mdl):
    return field_arr2vec(mdl.je_arr, mdl.order_cmpfirst), field_arr2vec(mdl.jm_arr, mdl.order_cmpfirst)


def create_stretched_dls(mdl):
    sdl = create_stretched_dl(mdl.wpml, mdl.grid, mdl.Npml, mdl.pml)
    ge = [ft2gt(EE, b) for b in mdl.boundft]
    gm = [ft2gt(HH, b) for b in mdl.boundft]
    K = len(mdl.grid)
    sdl_e = tuple(sdl[ge[w]][w] for w in range(K))
    sdl_m = tuple(sdl[gm[w]][w] for w in range(K))
    return sdl_e, sdl_m, tuple(1 / a for a in sdl_e), tuple(1 / a for a in sdl_m)


def _mu_or_none(mu):
    ident = np.zeros((3, 3))
    np.fill_diagonal(ident, 1.0)
    if not mu.any() or np.array_equal(mu, np.broadcast_to(ident, mu.shape)):
        return None          # zeros (never assigned) or identity: mu == 1
    return mu


class ParamOp:
    """What the reference's create_paramop turns into a sparse matrix (model.jl:152-155): the material array plus the
    averaging inputs.  Here it stays a description - the GPU kernels apply it matrix-free.  Holds a REFERENCE to the
    model's array (no copy of a multi-GB tensor), so build the operator before editing the model again."""

    _tokens = itertools.count(1)

    def __init__(self, kind, arr, geom):
        self.kind, self.arr, self.geom = kind, arr, geom
        # every create_paramops call yields new material operators (the reference builds new sparse matrices,
        # model.jl:152-155): operators cached on a shared Cs are keyed by this token, so Ps from a later call - after
        # the objects or the arrays changed - never resolve to a GPU operator that still holds the old material
        self.token = next(ParamOp._tokens)


class CurlOp:
    """What the reference's create_curl turns into a sparse matrix (model.jl:171-172): which curl (Ce: E->H, Cm: H->E)
    and the 1-D inputs (stretched dl's, Bloch flags and phases, boundft, DOF order)."""

    def __init__(self, kind, geom):
        self.kind, self.geom = kind, geom


class _Geom:
    """the model settings both create_paramops and create_curls read (model.jl:141-175), taken at call time"""

    def __init__(self, mdl):
        self.N, self.isbloch = mdl.grid.N, mdl.grid.isbloch
        self.sdl_e, self.sdl_m, _, _ = create_stretched_dls(mdl)
        self.e_mikL = create_e_mikL(mdl)
        self.boundft, self.order_cmpfirst = tuple(mdl.boundft), mdl.order_cmpfirst
        self.cmp_s, self.cmp_e, self.cmp_m = mdl.cmp_s, mdl.cmp_e, mdl.cmp_m
        self.ops = {}            # operators built from this description, keyed by (ft, w, options)

    def same(self, o):
        return (self.N == o.N and self.isbloch == o.isbloch and self.boundft == o.boundft
                and self.order_cmpfirst == o.order_cmpfirst and np.array_equal(self.e_mikL, o.e_mikL)
                and (self.cmp_s, self.cmp_e, self.cmp_m) == (o.cmp_s, o.cmp_e, o.cmp_m)
                and all(np.array_equal(a, b) for a, b in zip(self.sdl_e + self.sdl_m, o.sdl_e + o.sdl_m)))


def create_paramops(mdl, device=-1, device_materials=False):
    """(Peps, Pmu) - model.jl:141-158.  As in the reference (:143) the material arrays are first computed from the
    objects added with add_obj (calc_matparams, GPU); a model without objects keeps arrays that were filled directly.
    device_materials=True (models whose objects all have mu = 1): mdl.eps_arr is NOT filled - Peps carries the objects
    and the operator rasterises and smooths its own z-slab on the device (fdfd_set_eps_objects), so no (Nx,Ny,Nz,3,3)
    host array is ever built."""
    objs = getattr(mdl, "oind2shp", None)
    g = _Geom(mdl)
    if objs and device_materials and len(mdl.grid) < 3:
        raise ValueError("device_materials needs a 3-D model")
    if objs and device_materials:
        if not (len(mdl.muind2mu) == 1 and np.array_equal(mdl.muind2mu[0], np.eye(3))):
            raise ValueError("device_materials needs mu = 1 for every object")
        Pe = ParamOp("eps", None, g)
        Pe.objects = (mdl.grid.lg_prim, list(mdl.oind2shp), list(mdl.oind2epsind), list(mdl.epsind2eps))
        mu = np.zeros(mdl.grid.N + (3, 3), np.complex128)
        for v in range(3):
            mu[..., v, v] = 1.0
        return Pe, ParamOp("mu", mu, g)
    if objs:
        from .shapes import calc_matparams
        calc_matparams(mdl, device=device)
    return ParamOp("eps", mdl.eps_arr, g), ParamOp("mu", mdl.mu_arr, g)


def create_curls(mdl):
    """(Ce, Cm) - model.jl:160-175."""
    g = _Geom(mdl)
    return CurlOp("Ce", g), CurlOp("Cm", g)


def _build(ft, w, Ps, Cs, device=-1, rank=0, nranks=1, kernel=0, weighted_out_avg=False, ngpu=None, devices=None):
    if ft not in (EE, HH):
        raise ValueError(f"ft = {ft} is unsupported.")          # model.jl:242
    Pe, Pm = Ps
    Ce, Cm = Cs
    if not (isinstance(Pe, ParamOp) and isinstance(Pm, ParamOp) and isinstance(Ce, CurlOp) and isinstance(Cm, CurlOp)):
        raise TypeError("Ps / Cs must come from create_paramops / create_curls")
    g = Ce.geom
    if not g.same(Pe.geom):
        raise ValueError("create_paramops and create_curls were called on different model settings")
    if ngpu is not None:
        device = ("multi", int(ngpu), None if devices is None else tuple(int(v) for v in devices))
    key = (ft, complex(w), device, rank, nranks, kernel, weighted_out_avg, Pe.token, Pm.token)
    A = g.ops.get(key)
    if A is not None and not A.closed:
        return A
    for k in [k for k in g.ops if k[:7] == key[:7]]:     # same settings, older material: no longer reachable through
        del g.ops[k]                                     # the cache (the caller's own reference keeps it alive)
    if len(g.N) < 3:
        # ModelTE / ModelTM / ModelTEM: the 3-D handle that is one periodic cell thick along the missing axes
        from .reduced import ReducedOperator, embed_geometry, embed_param
        if nranks != 1:
            raise ValueError("z-slabs need a 3-D model")
        N3, bl3, se3, sm3, ph3, bft3 = embed_geometry(g)
        eps3 = embed_param(Pe.arr, N3, g.cmp_e)
        if eps3 is None:
            eps3 = np.broadcast_to(np.eye(3, dtype=np.complex128), N3 + (3, 3)).copy()
        A3 = FdfdOperator(N3, bl3, se3, sm3, w, eps3, embed_param(Pm.arr, N3, g.cmp_m), ph3, boundft=bft3, ft=ft,
                          order_cmpfirst=True, device=device, kernel=kernel, weighted_out_avg=weighted_out_avg)
        A = ReducedOperator(A3, int(np.prod(g.N)), g.cmp_e, g.cmp_m, ft, g.order_cmpfirst)
        g.ops[key] = A
        return A
    mu = _mu_or_none(Pm.arr)
    objects = getattr(Pe, "objects", None)
    if ngpu is not None:
        # one value over ngpu devices of this box (fdfd_multi_*): full-grid arrays in, the library cuts the z-slabs
        from .operator import MultiGpuOperator
        if nranks != 1 or rank != 0:
            raise ValueError("ngpu (one process, several GPUs) and rank / nranks (one process per GPU) exclude each other")
        A = MultiGpuOperator(g.N, g.isbloch, g.sdl_e, g.sdl_m, w, None if objects else Pe.arr, mu, g.e_mikL, boundft=g.boundft,
                             ft=ft, order_cmpfirst=g.order_cmpfirst, ngpu=ngpu, devices=devices, kernel=kernel,
                             weighted_out_avg=weighted_out_avg)
        if objects:
            lprim, shapes, pinds, params = objects
            A.set_eps_objects(lprim, shapes, pinds, params, boundft=g.boundft)
        g.ops[key] = A
        return A
    from .operator import partition
    k0, k1 = partition(g.N[2], nranks, rank)
    A = FdfdOperator(g.N, g.isbloch, g.sdl_e, g.sdl_m, w, None if objects else Pe.arr[:, :, k0:k1],
                     None if mu is None else mu[:, :, k0:k1], g.e_mikL, boundft=g.boundft, ft=ft,
                     order_cmpfirst=g.order_cmpfirst, device=device, rank=rank, nranks=nranks, kernel=kernel,
                     weighted_out_avg=weighted_out_avg)
    if objects:
        lprim, shapes, pinds, params = objects
        A.set_eps_objects(lprim, shapes, pinds, params, boundft=g.boundft)
    g.ops[key] = A
    return A


def _as_ops(third, fourth):
    """accept both call shapes: (..., mdl) and the reference's (..., Ps, Cs)"""
    if isinstance(third, Model):
        return create_paramops(third), create_curls(third)
    return third, fourth


def create_A(ft, w, Ps, Cs=None, **kw):
    """create_A(ft, ω, Ps, Cs) (model.jl:222-246) - or create_A(ft, ω, mdl).  Returns the GPU operator (an
    FdfdOperator: supports `A @ x`, `A.solve(b)`), not a SparseMatrixCSC; nothing is assembled.  Keyword options:
    device, rank, nranks (z-slabs: this rank's slab of eps/mu is taken from the full model arrays), kernel,
    weighted_out_avg; ngpu=N (and optionally devices=[...]) returns ONE operator over N GPUs of this box that takes
    full-grid host vectors (MultiGpuOperator, fdfd_multi_*)."""
    Ps, Cs = _as_ops(Ps, Cs)
    return _build(ft, w, Ps, Cs, **kw)


def create_b(ft, w, Ps, Cs=None, js=None, **kw):
    """create_b(ft, ω, Ps, Cs, js) (model.jl:248-274), evaluated on the GPU (fdfd_create_b): EE: b = -Cm(Pmu\\jm) - iω je,
    HH: b = Ce(Peps\\je) - iω jm.  Also accepts create_b(ft, ω, A, js) with an operator built by create_A."""
    if isinstance(Ps, _OPS):
        A, js = Ps, (Cs if js is None else js)
    else:
        Ps, Cs = _as_ops(Ps, Cs)
        A = _build(ft, w, Ps, Cs, **kw)
    if ft not in (EE, HH):
        raise ValueError(f"ft = {ft} is unsupported.")
    if A.ft != ft:
        raise ValueError("create_b: operator was built for the other formulation")
    je, jm = js
    other = jm if ft == EE else je        # the current that goes through the curl; skipped when identically zero
    if ft == EE:
        return A.create_b(je, jm if np.any(other) else None)
    return A.create_b(je, jm)


def create_linsys(ft, w, Ps, Cs=None, js=None, **kw):
    """create_linsys(ft, ω, Ps, Cs, js) (model.jl:209-220) - or create_linsys(ft, ω, mdl) (sources from the model)."""
    if isinstance(Ps, Model):
        js = create_srcs(Ps) if js is None else js
    Ps, Cs = _as_ops(Ps, Cs)
    A = _build(ft, w, Ps, Cs, **kw)
    return A, create_b(ft, w, A, js)


def _post_operator(w, third, fourth, **kw):
    if isinstance(third, _OPS):
        return third
    Ps, Cs = _as_ops(third, fourth)
    g = Cs[0].geom
    for key, A in g.ops.items():           # reuse the operator create_A built for this ω (either formulation works)
        if key[1] == complex(w) and key[7:] == (Ps[0].token, Ps[1].token) and not A.closed:
            return A
    return _build(EE, w, Ps, Cs, **kw)


def h_from_e(e, w, Ps, Cs=None, js=None, **kw):
    """h_from_e(e, ω, Ps, Cs, js) (model.jl:276-279) = (i/ω) Pmu \\ (Ce e + jm); also h_from_e(e, ω, A, jm)."""
    if isinstance(Ps, _OPS):
        jm = Cs if js is None else js
    else:
        jm = None if js is None else js[1]
    A = _post_operator(w, Ps, Cs, **kw)
    return A.h_from_e(e, jm if (jm is not None and np.any(jm)) else None)


def e_from_h(h, w, Ps, Cs=None, js=None, **kw):
    """e_from_h(h, ω, Ps, Cs, js) (model.jl:281-284) = (-i/ω) Peps \\ (Cm h - je); also e_from_h(h, ω, A, je)."""
    if isinstance(Ps, _OPS):
        je = Cs if js is None else js
    else:
        je = None if js is None else js[0]
    A = _post_operator(w, Ps, Cs, **kw)
    return A.e_from_h(h, je if (je is not None and np.any(je)) else None)


def create_Mcs(mdl_or_A, **kw):
    """create_Mcs(mdl) (model.jl:287-306): two callables (Mc_e, Mc_m) interpolating E / H to the voxel corners
    (the reference returns two sparse matrices; apply these like `Mc_e(e)`).  Also accepts an operator."""
    A = mdl_or_A if isinstance(mdl_or_A, _OPS) else _build(EE, 0.0, *(_as_ops(mdl_or_A, None)), **kw)
    return (lambda e: A.interp_corners(e, "E")), (lambda h: A.interp_corners(h, "H"))


def solve(A, b, **kw):
    return A.solve(b, **kw)
