"""Host-side mirror of the reference's model API for the hot path (reference src/model/model.jl).

Same names and argument meaning as the reference (ASCII spellings: ω→w, ε→eps, μ→mu, ∆→d, ₑ→e, ₘ→m):
    ModelFull(grid)                     full.jl:7-13
    set_wpml / set_boundft / set_Npml / set_kbloch      model.jl:86-89
    create_e_mikL                       model.jl:91
    clear_srcs / add_srce / add_srcm / create_srcs      model.jl:177-207
    create_stretched_dls                model.jl:122-139
    create_A / create_b / create_linsys model.jl:209-274   (GPU: returns an FdfdOperator, not a CSC)
    h_from_e                            model.jl:276-279
    solve                               (absent in the reference: `A \\ b` left to the user)
What changes for the user: `Ps = create_paramops(mdl); Cs = create_curls(mdl); A,b = create_linsys(EE,ω,Ps,Cs,js);
e = A\\b` becomes `A,b = create_linsys(EE, ω, mdl); e,info = solve(A,b)`; the operator is never assembled.
Geometry rasterisation (calc_matparams!, full.jl:16-70) is out of scope: fill mdl.eps_arr / mdl.mu_arr directly.
"""
import numpy as np

from .grid import EE, HH, PRIM, DUAL, Grid, PMLParam, create_stretched_dl, ft2gt
from .operator import FdfdOperator
from .sources import Source


class Model:
    def __init__(self, grid):
        K = len(grid)
        self.wpml = 0.0                                   # ωpml (model.jl:37)
        self.grid = grid
        self.boundft = (EE,) * K                          # model.jl:46
        self.Npml = ((0,) * K, (0,) * K)                  # model.jl:47
        self.kbloch = (0.0,) * K                          # model.jl:48
        self.eps_arr = np.zeros(grid.N + (3, 3), np.complex128)   # create_param_array (model.jl:51-52)
        self.mu_arr = np.zeros(grid.N + (3, 3), np.complex128)
        self.je_arr = np.zeros(grid.N + (3,), np.complex128)      # create_field_array (model.jl:55-56)
        self.jm_arr = np.zeros(grid.N + (3,), np.complex128)
        self.order_cmpfirst = True                        # model.jl:72
        self.pml = PMLParam()

    def size(self, ft):
        return ((3,) + self.grid.N) if self.order_cmpfirst else (self.grid.N + (3,))   # model.jl:75-81

    def length(self, ft):
        return int(np.prod(self.size(ft)))


def ModelFull(grid):
    if len(grid) != 3:
        raise ValueError("ModelFull needs a 3-D grid")
    return Model(grid)


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
    F = np.asarray(F)
    axes = (2, 1, 0, 3) if order_cmpfirst else (3, 2, 1, 0)
    return np.ascontiguousarray(F.transpose(axes)).reshape(-1)


def field_vec2arr(v, N, order_cmpfirst=True):
    Nx, Ny, Nz = N
    if order_cmpfirst:
        return np.asarray(v).reshape(Nz, Ny, Nx, 3).transpose(2, 1, 0, 3)
    return np.asarray(v).reshape(3, Nz, Ny, Nx).transpose(3, 2, 1, 0)


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


def _mu_or_none(mdl):
    mu = mdl.mu_arr
    ident = np.zeros((3, 3))
    np.fill_diagonal(ident, 1.0)
    if not mu.any() or np.array_equal(mu, np.broadcast_to(ident, mu.shape)):
        return None          # zeros (never assigned) or identity: mu == 1
    return mu


def create_A(ft, w, mdl, device=-1, rank=0, nranks=1, kernel=0, weighted_out_avg=False):
    """GPU stand-in for create_A(ft, ω, create_paramops(mdl), create_curls(mdl)) (model.jl:141-175,225-246).
    With nranks > 1 this rank's z-slab of eps/mu is taken from the full model arrays."""
    if ft not in (EE, HH):
        raise ValueError(f"ft = {ft} is unsupported.")
    sdl_e, sdl_m, _, _ = create_stretched_dls(mdl)
    from .operator import partition
    k0, k1 = partition(mdl.grid.N[2], nranks, rank)
    mu = _mu_or_none(mdl)
    return FdfdOperator(mdl.grid.N, mdl.grid.isbloch, sdl_e, sdl_m, w, mdl.eps_arr[:, :, k0:k1],
                        None if mu is None else mu[:, :, k0:k1], create_e_mikL(mdl),
                        boundft=mdl.boundft, ft=ft, order_cmpfirst=mdl.order_cmpfirst, device=device,
                        rank=rank, nranks=nranks, kernel=kernel, weighted_out_avg=weighted_out_avg)


def create_b(ft, w, A, js):
    """create_b (model.jl:251-274), EE branch, evaluated on the GPU through fdfd_create_b."""
    if ft != EE:
        raise ValueError(f"ft = {ft} is unsupported.")
    je, jm = js
    return A.create_b(je, jm if np.any(jm) else None)


def create_linsys(ft, w, mdl, **kw):
    A = create_A(ft, w, mdl, **kw)
    return A, create_b(ft, w, A, create_srcs(mdl))


def h_from_e(e, w, A, jm=None):
    return A.h_from_e(e, jm)


def e_from_h(h, w, A, je=None):
    """e_from_h (model.jl:281-284)"""
    return A.e_from_h(h, je)


def create_Mcs(A):
    """create_Mcs (model.jl:287-306): returns two callables (Mc_e, Mc_m) interpolating E / H to the voxel corners."""
    return (lambda e: A.interp_corners(e, "E")), (lambda h: A.interp_corners(h, "H"))


def solve(A, b, **kw):
    return A.solve(b, **kw)
