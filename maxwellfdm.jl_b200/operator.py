"""GPU operator object: what `create_A` returns in the reference (src/model/model.jl:225-246) -
something that supports `A * x`, `mul!(y, A, x)` and a solve - backed by libfdfd_b200.so."""
import ctypes as C

import numpy as np

from . import _lib as L


def _ptr(a):
    """Raw address of a numpy array, a torch tensor, an int, or None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return a.ctypes.data


def _where(a):
    return L.DEVICE if (hasattr(a, "is_cuda") and a.is_cuda) else L.HOST


def _c128(z):
    z = complex(z)
    return L.c128(z.real, z.imag)


def _matparams_desc(N, isbloch, lprim, shapes, pinds, params, boundft, field_ortho_shape):
    """fdfd_matparams_desc for an object list (+ the buffers it points into, which the caller must keep alive)"""
    from .shapes import _as_tensor
    sh = (L.Shape * len(shapes))()
    for s, shp, pi in zip(sh, shapes, pinds):
        s.kind, s.axis, s.pind = shp.kind, shp.axis, int(pi)
        s.c[:] = shp.c
        s.r[:] = shp.r
    prm = np.ascontiguousarray(np.stack([_as_tensor(P) for P in params]).reshape(-1, 9))
    lp = [np.ascontiguousarray(a, dtype=np.float64) for a in lprim]
    d = L.MatParamsDesc()
    d.N[:] = N
    d.isbloch[:] = [1 if b else 0 for b in isbloch]
    d.boundft_is_E[:] = [1 if str(b).upper().startswith("E") or b == 0 else 0 for b in boundft]
    d.field_type = L.FT_EE
    d.field_ortho_shape = 1 if field_ortho_shape else 0
    d.lprim[:] = [a.ctypes.data for a in lp]
    d.nshape, d.nparam = len(shapes), prm.shape[0]
    d.shapes = C.cast(sh, C.c_void_p).value
    d.params = prm.ctypes.data
    return d, (sh, prm, lp)


class FdfdOperator:
    """Matrix-free A = C2 q C1 - w^2 P on one z-slab of the grid.

    Parameters mirror the quantities the reference passes to create_curl / create_paramop
    (model.jl:152-155,171-172): stretched cell sizes, Bloch flags and phases, eps/mu arrays.
    eps / mu are numpy arrays indexed [i,j,k,v,u] of THIS slab (shape (Nx,Ny,nzl,3,3))."""

    def __init__(self, N, isbloch, sdl_e, sdl_m, omega, eps, mu=None, e_mikL=(1, 1, 1),
                 boundft=("E", "E", "E"), ft="E", order_cmpfirst=True, device=-1, rank=0, nranks=1,
                 weighted_out_avg=False, kernel=L.KERNEL_AUTO, eps_has_offdiag=None):
        self._h = None
        lib = L.lib()
        d = L.Desc()
        d.N[:] = [int(n) for n in N]
        d.isbloch[:] = [1 if b else 0 for b in isbloch]
        d.boundft_is_E[:] = [1 if str(b).upper().startswith("E") or b == 0 else 0 for b in boundft]
        d.order_cmpfirst = 1 if order_cmpfirst else 0
        d.field_type = L.FT_EE if (str(ft).upper().startswith("E") or ft == 0) else L.FT_HH
        d.device, d.rank, d.nranks = int(device), int(rank), int(nranks)
        d.weighted_out_avg = 1 if weighted_out_avg else 0
        d.kernel = int(kernel)
        h = C.c_void_p()
        L.check(lib.fdfd_create(C.byref(h), C.byref(d)))
        self._h = h
        self.ft = 0 if d.field_type == L.FT_EE else 1        # EE / HH as in grid.py
        self.N = tuple(int(n) for n in N)
        self._isbloch = tuple(bool(b) for b in isbloch)
        self.order_cmpfirst = bool(order_cmpfirst)
        k0, k1 = C.c_int64(), C.c_int64()
        L.check(lib.fdfd_slab_range(h, C.byref(k0), C.byref(k1)), h)
        self.k0, self.k1 = k0.value, k1.value
        self.nzl = self.k1 - self.k0
        self.n = 3 * self.N[0] * self.N[1] * self.nzl          # local DOFs
        self.shape = (self.n, self.n)
        self.set_coeffs(sdl_e, sdl_m)
        self.set_bloch(e_mikL)
        self.set_omega(omega)
        if eps is not None:
            self.set_eps(eps, eps_has_offdiag)
        self.set_mu(mu)

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if self._h is not None:
            L.lib().fdfd_destroy(self._h)
            self._h = None

    @property
    def closed(self):
        return self._h is None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- inputs -----------------------------------------------------------------------------
    def set_coeffs(self, sdl_e, sdl_m):
        keep = [np.ascontiguousarray(a, dtype=np.complex128) for a in list(sdl_e) + list(sdl_m)]
        for a, n in zip(keep, self.N * 2):
            if a.shape != (n,):
                raise ValueError("sdl arrays must have the global axis length")
        pe = (C.c_void_p * 3)(*[a.ctypes.data for a in keep[:3]])
        pm = (C.c_void_p * 3)(*[a.ctypes.data for a in keep[3:]])
        L.check(L.lib().fdfd_set_coeffs(self._h, pe, pm), self._h)

    def set_bloch(self, e_mikL):
        ph = np.ascontiguousarray(e_mikL, dtype=np.complex128)
        L.check(L.lib().fdfd_set_bloch(self._h, ph.ctypes.data), self._h)

    def set_omega(self, omega):
        self.omega = complex(omega)
        L.check(L.lib().fdfd_set_omega(self._h, _c128(omega)), self._h)

    @staticmethod
    def _julia_layout(p):
        """(Nx,Ny,nzl,3,3)-indexed numpy array -> memory of a Julia column-major array of that size."""
        p = np.asarray(p, dtype=np.complex128)
        return np.ascontiguousarray(p.transpose(4, 3, 2, 1, 0))

    def set_eps(self, eps, has_offdiag=None):
        """eps indexed [i,j,k,v,u] (shape (Nx,Ny,nzl,3,3)), or - to avoid a transposing copy of a large array - a
        C-contiguous complex128 array of shape (3,3,nzl,Ny,Nx) indexed [u,v,k,j,i], which IS the memory of the Julia
        column-major (Nx,Ny,nzl,3,3) array."""
        eps = np.asarray(eps)
        # (a 3x3 grid in x-y makes the two shapes coincide: such an array is read as [i,j,k,v,u])
        if eps.shape == (3, 3, self.nzl, self.N[1], self.N[0]) and eps.dtype == np.complex128 and eps.flags.c_contiguous \
                and (self.N[1], self.N[0]) != (3, 3):
            buf = eps
        elif eps.shape != (self.N[0], self.N[1], self.nzl, 3, 3):
            raise ValueError(f"eps must have shape (Nx,Ny,nzl,3,3) = {(self.N[0], self.N[1], self.nzl, 3, 3)}")
        else:
            buf = self._julia_layout(eps)
        if has_offdiag is None:
            has_offdiag = True  # the library scans and drops the flag if every off-diagonal entry is 0
        L.check(L.lib().fdfd_set_eps(self._h, buf.ctypes.data, 1 if has_offdiag else 0), self._h)

    def set_eps_objects(self, lprim, shapes, pinds, params, boundft=("E", "E", "E"), isbloch=None, field_ortho_shape=False):
        """eps straight from objects (shapes.py): this rank's slab is rasterised and subpixel-smoothed on the device,
        directly into the operator's material arrays - no (Nx,Ny,Nz,3,3) host array.  lprim: the Grid's ghosted primal
        planes; shapes / pinds / params as for calc_matparams_array; boundft / isbloch must be the handle's."""
        d, keep = _matparams_desc(self.N, self._isbloch if isbloch is None else isbloch, lprim, shapes, pinds, params, boundft,
                                  field_ortho_shape)
        L.check(L.lib().fdfd_set_eps_objects(self._h, C.byref(d)), self._h)

    def set_mu(self, mu):
        if mu is None:
            L.check(L.lib().fdfd_set_mu(self._h, None), self._h)
            return
        mu = np.asarray(mu)
        if mu.shape != (self.N[0], self.N[1], self.nzl, 3, 3):
            raise ValueError("mu must have shape (Nx,Ny,nzl,3,3)")
        buf = self._julia_layout(mu)
        L.check(L.lib().fdfd_set_mu(self._h, buf.ctypes.data), self._h)

    def comm_init(self, unique_id: bytes):
        L.check(L.lib().fdfd_comm_init(self._h, unique_id), self._h)

    # -- operator ---------------------------------------------------------------------------
    def _out_like(self, x):
        if hasattr(x, "is_cuda"):
            import torch
            return torch.empty_like(x)
        return np.empty(self.n, dtype=np.complex128)

    def _chk_vec(self, x, name):
        if hasattr(x, "is_cuda"):
            import torch
            if x.dtype != torch.complex128 or x.numel() != self.n or not x.is_contiguous():
                raise ValueError(f"{name} must be a contiguous complex128 tensor of {self.n} elements")
            if x.is_cuda:
                torch.cuda.current_stream(x.device).synchronize()
            return x
        x = np.ascontiguousarray(x, dtype=np.complex128)
        if x.shape != (self.n,):
            raise ValueError(f"{name} must have {self.n} elements")
        return x

    def _chk_out(self, y, like, name):
        """an OUTPUT buffer is written in place by the library: it must already be what the C ABI expects, and live
        where the input lives (the `where` argument covers both)"""
        if hasattr(like, "is_cuda"):
            import torch
            if not hasattr(y, "is_cuda") or y.dtype != torch.complex128 or y.numel() != self.n or not y.is_contiguous():
                raise ValueError(f"{name} must be a contiguous complex128 tensor of {self.n} elements")
            if y.is_cuda != like.is_cuda:
                raise ValueError(f"{name} and its input must both be host or both be device buffers")
            return y
        if hasattr(y, "is_cuda") or not isinstance(y, np.ndarray) or y.dtype != np.complex128 or y.shape != (self.n,) \
                or not y.flags.c_contiguous or not y.flags.writeable:
            raise ValueError(f"{name} must be a writeable contiguous complex128 numpy array of {self.n} elements "
                             f"(a host buffer, like its input)")
        return y

    def mul(self, y, x, transpose=False):
        """mul!(y, A, x)"""
        x = self._chk_vec(x, "x")
        y = self._chk_out(y, x, "y")
        f = L.lib().fdfd_apply_transpose if transpose else L.lib().fdfd_apply
        L.check(f(self._h, _ptr(x), _ptr(y), _where(x)), self._h)
        return y

    def __matmul__(self, x):
        x = self._chk_vec(x, "x")
        return self.mul(self._out_like(x), x)

    __mul__ = __matmul__

    def rmatvec_T(self, x):
        x = self._chk_vec(x, "x")
        return self.mul(self._out_like(x), x, transpose=True)

    def solve(self, b, x0=None, method="bicgstab", rtol=1e-8, maxit=10000, check_every=10, history=False):
        """x = A \\ b by BiCGSTAB or QMR.  Returns (x, info)."""
        b = self._chk_vec(b, "b")
        if x0 is None:
            if hasattr(b, "is_cuda"):
                import torch
                x = torch.zeros_like(b)
            else:
                x = np.zeros(self.n, dtype=np.complex128)
        else:
            x = self._chk_vec(x0, "x0")
            if hasattr(x, "is_cuda") != hasattr(b, "is_cuda") or (hasattr(x, "is_cuda") and x.is_cuda != b.is_cuda):
                raise ValueError("x0 and b must both be host or both be device buffers")
            x = x.clone() if hasattr(x, "clone") else x.copy()
        m = L.BICGSTAB if str(method).lower().startswith("bi") else L.QMR
        iters, relres = C.c_int(), C.c_double()
        hist = np.full(maxit + 1, np.nan) if history else None
        code = L.lib().fdfd_solve(self._h, m, _ptr(b), _ptr(x), _where(b), float(rtol), int(maxit),
                                  int(check_every), C.byref(iters), C.byref(relres),
                                  hist.ctypes.data if history else None)
        L.check(code, self._h, ok=(L.OK, L.ENOCONV))
        info = {"iters": iters.value, "relres": relres.value, "converged": code == L.OK}
        if history:
            info["history"] = hist[: iters.value + 1]
        return x, info

    def export_pattern(self, values=True):
        """(colptr, rowval, nzval) of the assembled A as Julia stores it (1-based Int64)."""
        nnz = C.c_int64(0)
        L.check(L.lib().fdfd_export_pattern(self._h, None, None, None, C.byref(nnz)), self._h)
        colptr = np.empty(self.n + 1, dtype=np.int64)
        rowval = np.empty(nnz.value, dtype=np.int64)
        nzval = np.empty(nnz.value, dtype=np.complex128) if values else None
        L.check(L.lib().fdfd_export_pattern(self._h, colptr.ctypes.data, rowval.ctypes.data,
                                            nzval.ctypes.data if values else None, C.byref(nnz)), self._h)
        return colptr, rowval, nzval

    def h_from_e(self, e, jm=None):
        e = self._chk_vec(e, "e")
        if jm is not None:
            jm = self._chk_vec(jm, "jm")
        h = self._out_like(e)
        L.check(L.lib().fdfd_h_from_e(self._h, _ptr(e), _ptr(jm), _ptr(h), _where(e)), self._h)
        return h

    def e_from_h(self, h, je=None):
        h = self._chk_vec(h, "h")
        if je is not None:
            je = self._chk_vec(je, "je")
        e = self._out_like(h)
        L.check(L.lib().fdfd_e_from_h(self._h, _ptr(h), _ptr(je), _ptr(e), _where(h)), self._h)
        return e

    def interp_corners(self, f, ft="E"):
        """Mc_e * f (ft='E') or Mc_m * f (ft='H'): fields interpolated to the voxel corners (create_Mcs)."""
        f = self._chk_vec(f, "f")
        out = self._out_like(f)
        which = L.FT_EE if (str(ft).upper().startswith("E") or ft == 0) else L.FT_HH
        L.check(L.lib().fdfd_interp_corners(self._h, which, _ptr(f), _ptr(out), _where(f)), self._h)
        return out

    def create_b(self, je, jm=None):
        je = self._chk_vec(je, "je")
        if jm is not None:
            jm = self._chk_vec(jm, "jm")
        b = self._out_like(je)
        L.check(L.lib().fdfd_create_b(self._h, _ptr(je), _ptr(jm), _ptr(b), _where(je)), self._h)
        return b

    # -- measurement ------------------------------------------------------------------------
    def bench_apply(self, x_dev, y_dev, warmup=3, iters=10, flush_l2=False):
        tot, mn = C.c_double(), C.c_double()
        L.check(L.lib().fdfd_bench_apply(self._h, _ptr(x_dev), _ptr(y_dev), warmup, iters, 1 if flush_l2 else 0,
                                         C.byref(tot), C.byref(mn)), self._h)
        return tot.value, mn.value

    def bench_solve(self, b_dev, x_dev, method="bicgstab", warmup=2, iters=20):
        m = L.BICGSTAB if str(method).lower().startswith("bi") else L.QMR
        tot = C.c_double()
        L.check(L.lib().fdfd_bench_solve(self._h, m, _ptr(b_dev), _ptr(x_dev), warmup, iters, C.byref(tot)), self._h)
        return tot.value

    def bench_halo(self, x_dev, warmup=3, iters=20):
        """(ms for `iters` halo exchanges of x_dev by themselves, bytes this rank sends per exchange)"""
        tot, nb = C.c_double(), C.c_uint64()
        L.check(L.lib().fdfd_bench_halo(self._h, _ptr(x_dev), warmup, iters, C.byref(tot), C.byref(nb)), self._h)
        return tot.value, int(nb.value)

    @property
    def halo_data_plane(self):
        """'none' (single slab), 'nccl' (grouped ncclSend / ncclRecv) or 'peer' (copy-engine exchange into IPC-mapped buffers)"""
        k = C.c_int()
        L.check(L.lib().fdfd_halo_data_plane(self._h, C.byref(k)), self._h)
        return ("none", "nccl", "peer")[k.value]

    @property
    def offdiag_fraction(self):
        f = C.c_double()
        L.check(L.lib().fdfd_offdiag_fraction(self._h, C.byref(f)), self._h)
        return f.value

    @property
    def mass_bytes_per_dof(self):
        """bytes per DOF the kernel streams for the diagonal material terms (16; 8 when the mass entries are real)"""
        f = C.c_double()
        L.check(L.lib().fdfd_mass_bytes_per_dof(self._h, C.byref(f)), self._h)
        return f.value

    @property
    def offdiag_bytes_per_dof(self):
        """bytes per DOF of the off-diagonal streams on a flagged block: 32, 16 (symmetric), 8 (symmetric and real, fused
        row-pair kernel), 0 (none)"""
        f = C.c_double()
        L.check(L.lib().fdfd_offdiag_bytes_per_dof(self._h, C.byref(f)), self._h)
        return f.value

    @property
    def offdiag_symmetric(self):
        """True when the off-diagonal mass entries are pointwise symmetric (stored once: 16 instead of 32 B/DOF)."""
        f = C.c_int()
        L.check(L.lib().fdfd_offdiag_symmetric(self._h, C.byref(f)), self._h)
        return bool(f.value)

    @property
    def launch_count(self):
        return int(L.lib().fdfd_launch_count(self._h))


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    L.check(L.lib().fdfd_comm_unique_id(buf))
    return buf.raw


def partition(Nz, nranks, rank):
    k0, k1 = C.c_int64(), C.c_int64()
    code = L.lib().fdfd_partition(int(Nz), int(nranks), int(rank), C.byref(k0), C.byref(k1))
    if code != L.OK:
        raise ValueError("bad partition arguments")
    return k0.value, k1.value


def halo_plan(nranks, rank, wrapz):
    """(up, dn) neighbour ranks of the z-slab halo exchange (-1: symmetry boundary, no message)."""
    up, dn = C.c_int32(), C.c_int32()
    if L.lib().fdfd_halo_plan(int(nranks), int(rank), 1 if wrapz else 0, C.byref(up), C.byref(dn)) != L.OK:
        raise ValueError("bad halo_plan arguments")
    return up.value, dn.value


class MultiGpuOperator:
    """ONE operator value over N GPUs of the box (fdfd_multi_*, csrc/multi.cpp): what a single-process host -
    a Julia session holding `A = create_A(...)`, model.jl:225-246 - uses instead of one FdfdOperator per process.  Takes
    FULL-GRID host arrays; the library splits them into z-slabs, one per device, one host thread each.

    `A @ x` moves every slab (with the two neighbour planes it needs) over that GPU's own PCIe link - no exchange between
    GPUs; `A.solve(b)` runs the slab Krylov loops with halos and inner products over NCCL."""

    def __init__(self, N, isbloch, sdl_e, sdl_m, omega, eps, mu=None, e_mikL=(1, 1, 1), boundft=("E", "E", "E"), ft="E",
                 order_cmpfirst=True, ngpu=1, devices=None, weighted_out_avg=False, kernel=L.KERNEL_AUTO, eps_has_offdiag=None):
        self._m = None
        lib = L.lib()
        d = L.Desc()
        d.N[:] = [int(n) for n in N]
        d.isbloch[:] = [1 if b else 0 for b in isbloch]
        d.boundft_is_E[:] = [1 if str(b).upper().startswith("E") or b == 0 else 0 for b in boundft]
        d.order_cmpfirst = 1 if order_cmpfirst else 0
        d.field_type = L.FT_EE if (str(ft).upper().startswith("E") or ft == 0) else L.FT_HH
        d.device, d.rank, d.nranks = -1, 0, 1
        d.weighted_out_avg = 1 if weighted_out_avg else 0
        d.kernel = int(kernel)
        devs = None if devices is None else (C.c_int32 * int(ngpu))(*[int(v) for v in devices])
        m = C.c_void_p()
        L.check_multi(lib.fdfd_multi_create(C.byref(m), C.byref(d), int(ngpu), devs))
        self._m = m
        self.ft = 0 if d.field_type == L.FT_EE else 1
        self.N = tuple(int(n) for n in N)
        self._isbloch = tuple(bool(b) for b in isbloch)
        self.ngpu = int(ngpu)
        self.n = 3 * self.N[0] * self.N[1] * self.N[2]
        self.shape = (self.n, self.n)
        self.order_cmpfirst = bool(order_cmpfirst)
        keep = [np.ascontiguousarray(a, dtype=np.complex128) for a in list(sdl_e) + list(sdl_m)]
        pe = (C.c_void_p * 3)(*[a.ctypes.data for a in keep[:3]])
        pm = (C.c_void_p * 3)(*[a.ctypes.data for a in keep[3:]])
        L.check_multi(lib.fdfd_multi_set_coeffs(m, pe, pm), m)
        ph = np.ascontiguousarray(e_mikL, dtype=np.complex128)
        L.check_multi(lib.fdfd_multi_set_bloch(m, ph.ctypes.data), m)
        self.omega = complex(omega)
        L.check_multi(lib.fdfd_multi_set_omega(m, _c128(omega)), m)
        if eps is not None:
            buf = self._material(eps, "eps")
            L.check_multi(lib.fdfd_multi_set_eps(m, buf.ctypes.data, 0 if eps_has_offdiag is False else 1), m)
        if mu is not None:
            buf = self._material(mu, "mu")
            L.check_multi(lib.fdfd_multi_set_mu(m, buf.ctypes.data), m)

    def _material(self, a, name):
        a = np.asarray(a)
        Nx, Ny, Nz = self.N
        if a.shape == (3, 3, Nz, Ny, Nx) and a.dtype == np.complex128 and a.flags.c_contiguous and (Ny, Nx) != (3, 3):
            return a                      # already the memory of the Julia column-major (Nx,Ny,Nz,3,3) array
        if a.shape != (Nx, Ny, Nz, 3, 3):
            raise ValueError(f"{name} must have shape (Nx,Ny,Nz,3,3) = {(Nx, Ny, Nz, 3, 3)} (the whole grid)")
        return FdfdOperator._julia_layout(a)

    def set_eps_objects(self, lprim, shapes, pinds, params, boundft=("E", "E", "E"), field_ortho_shape=False):
        """eps straight from objects: every slab rasterises and smooths its own planes on its device (fdfd_multi_set_eps_objects)"""
        d, keep = _matparams_desc(self.N, self._isbloch, lprim, shapes, pinds, params, boundft, field_ortho_shape)
        L.check_multi(L.lib().fdfd_multi_set_eps_objects(self._m, C.byref(d)), self._m)

    def close(self):
        if self._m is not None:
            L.lib().fdfd_multi_destroy(self._m)
            self._m = None

    @property
    def closed(self):
        return self._m is None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _vec(self, x, name, out=False):
        if hasattr(x, "is_cuda"):
            if x.is_cuda:
                raise ValueError(f"{name}: a multi-GPU operator takes full-grid HOST vectors")
            x = x.numpy()
        if out:
            if not isinstance(x, np.ndarray) or x.dtype != np.complex128 or x.shape != (self.n,) or not x.flags.c_contiguous \
                    or not x.flags.writeable:
                raise ValueError(f"{name} must be a writeable contiguous complex128 host array of {self.n} elements")
            return x
        x = np.ascontiguousarray(x, dtype=np.complex128)
        if x.shape != (self.n,):
            raise ValueError(f"{name} must have {self.n} elements")
        return x

    def slab(self, r):
        """(raw slab handle, k0, k1) of slab r - for the measurement entry points of the slab API"""
        h, k0, k1 = C.c_void_p(), C.c_int64(), C.c_int64()
        L.check_multi(L.lib().fdfd_multi_slab(self._m, int(r), C.byref(h), C.byref(k0), C.byref(k1)), self._m)
        return h, k0.value, k1.value

    def mul(self, y, x, transpose=False):
        """mul!(y, A, x) on full-grid host vectors"""
        x = self._vec(x, "x")
        yv = self._vec(y, "y", out=True)
        f = L.lib().fdfd_multi_apply_transpose if transpose else L.lib().fdfd_multi_apply
        L.check_multi(f(self._m, x.ctypes.data, yv.ctypes.data), self._m)
        return y

    def __matmul__(self, x):
        return self.mul(np.empty(self.n, dtype=np.complex128), x)

    __mul__ = __matmul__

    def solve(self, b, x0=None, method="bicgstab", rtol=1e-8, maxit=10000, check_every=10, history=False, out=None):
        """x = A \\ b.  Returns (x, info).  `out`: result buffer (e.g. pinned memory); it also carries the initial guess."""
        b = self._vec(b, "b")
        if out is not None:
            x = self._vec(out, "out", out=True)
            if x0 is not None:
                x[:] = self._vec(x0, "x0")
        else:
            x = np.zeros(self.n, dtype=np.complex128) if x0 is None else self._vec(x0, "x0").copy()
        m = L.BICGSTAB if str(method).lower().startswith("bi") else L.QMR
        iters, relres = C.c_int(), C.c_double()
        hist = np.full(maxit + 1, np.nan) if history else None
        code = L.lib().fdfd_multi_solve(self._m, m, b.ctypes.data, x.ctypes.data, float(rtol), int(maxit), int(check_every),
                                        C.byref(iters), C.byref(relres), hist.ctypes.data if history else None)
        L.check_multi(code, self._m, ok=(L.OK, L.ENOCONV))
        info = {"iters": iters.value, "relres": relres.value, "converged": code == L.OK}
        if history:
            info["history"] = hist[: iters.value + 1]
        return x, info

    def _vec_op(self, fn, a, b_or_none):
        a = self._vec(a, "input")
        b = None if b_or_none is None else self._vec(b_or_none, "input")
        out = np.empty(self.n, dtype=np.complex128)
        L.check_multi(fn(self._m, a.ctypes.data, None if b is None else b.ctypes.data, out.ctypes.data), self._m)
        return out

    def create_b(self, je, jm=None):
        return self._vec_op(L.lib().fdfd_multi_create_b, je, jm)

    def h_from_e(self, e, jm=None):
        return self._vec_op(L.lib().fdfd_multi_h_from_e, e, jm)

    def e_from_h(self, h, je=None):
        return self._vec_op(L.lib().fdfd_multi_e_from_h, h, je)
