"""ctypes binding of libfdfd_b200.so (C ABI declared in include/fdfd_b200.h).

The product path has NO CPU fallback: if the shared library is missing this raises at import of the
binding, and if no CUDA device is present fdfd_create fails with FDFD_ECUDA.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FDFD_B200_LIB", os.path.join(HERE, "libfdfd_b200.so"))   # override: tuning experiments

OK, EINVAL, ECUDA, ENCCL, ENOMEM, ENOCONV, ESTATE = range(7)
HOST, DEVICE = 0, 1
BICGSTAB, QMR = 0, 1
FT_EE, FT_HH = 0, 1
KERNEL_AUTO, KERNEL_NAIVE, KERNEL_TILED = 0, 1, 2
SHAPE_BOX, SHAPE_BALL, SHAPE_CYLINDER = 0, 1, 2


class c128(C.Structure):
    _fields_ = [("re", C.c_double), ("im", C.c_double)]


class Desc(C.Structure):
    _fields_ = [("N", C.c_int64 * 3), ("isbloch", C.c_int32 * 3), ("boundft_is_E", C.c_int32 * 3),
                ("order_cmpfirst", C.c_int32), ("field_type", C.c_int32), ("device", C.c_int32),
                ("rank", C.c_int32), ("nranks", C.c_int32), ("weighted_out_avg", C.c_int32),
                ("kernel", C.c_int32)]


class Shape(C.Structure):
    _fields_ = [("kind", C.c_int32), ("axis", C.c_int32), ("pind", C.c_int32), ("reserved", C.c_int32),
                ("c", C.c_double * 3), ("r", C.c_double * 3)]


class MatParamsDesc(C.Structure):
    _fields_ = [("N", C.c_int64 * 3), ("isbloch", C.c_int32 * 3), ("boundft_is_E", C.c_int32 * 3),
                ("field_type", C.c_int32), ("field_ortho_shape", C.c_int32), ("lprim", C.c_void_p * 3),
                ("k0", C.c_int64), ("k1", C.c_int64), ("nshape", C.c_int32), ("nparam", C.c_int32),
                ("shapes", C.c_void_p), ("params", C.c_void_p), ("device", C.c_int32)]


class FdfdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"fdfd_b200 error {code}: {msg}")
        self.code = code


_lib = None

# name -> (restype, argtypes); every symbol include/fdfd_b200.h declares
P = C.c_void_p
SYMBOLS = {
    "fdfd_version": (C.c_char_p, []),
    "fdfd_create": (C.c_int, [C.POINTER(P), C.POINTER(Desc)]),
    "fdfd_destroy": (C.c_int, [P]),
    "fdfd_last_error": (C.c_char_p, [P]),
    "fdfd_slab_range": (C.c_int, [P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "fdfd_partition": (C.c_int, [C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "fdfd_halo_plan": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "fdfd_set_coeffs": (C.c_int, [P, C.POINTER(P), C.POINTER(P)]),
    "fdfd_set_bloch": (C.c_int, [P, P]),
    "fdfd_set_omega": (C.c_int, [P, c128]),
    "fdfd_set_eps": (C.c_int, [P, P, C.c_int]),
    "fdfd_set_mu": (C.c_int, [P, P]),
    "fdfd_apply": (C.c_int, [P, P, P, C.c_int]),
    "fdfd_apply_transpose": (C.c_int, [P, P, P, C.c_int]),
    "fdfd_apply_host_halos": (C.c_int, [P, P, P, P, P, C.c_int]),
    "fdfd_solve": (C.c_int, [P, C.c_int, P, P, C.c_int, C.c_double, C.c_int, C.c_int,
                             C.POINTER(C.c_int), C.POINTER(C.c_double), P]),
    "fdfd_export_pattern": (C.c_int, [P, P, P, P, C.POINTER(C.c_int64)]),
    "fdfd_h_from_e": (C.c_int, [P, P, P, P, C.c_int]),
    "fdfd_e_from_h": (C.c_int, [P, P, P, P, C.c_int]),
    "fdfd_interp_corners": (C.c_int, [P, C.c_int, P, P, C.c_int]),
    "fdfd_create_b": (C.c_int, [P, P, P, P, C.c_int]),
    "fdfd_calc_matparams": (C.c_int, [P, P, C.c_int]),
    "fdfd_set_eps_objects": (C.c_int, [P, P]),
    "fdfd_comm_unique_id": (C.c_int, [C.c_char_p]),
    "fdfd_comm_init": (C.c_int, [P, C.c_char_p]),
    "fdfd_bench_apply": (C.c_int, [P, P, P, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "fdfd_bench_solve": (C.c_int, [P, C.c_int, P, P, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "fdfd_mass_bytes_per_dof": (C.c_int, [P, C.POINTER(C.c_double)]),
    "fdfd_bench_halo": (C.c_int, [P, P, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "fdfd_halo_data_plane": (C.c_int, [P, C.POINTER(C.c_int)]),
    "fdfd_set_shared_process": (C.c_int, [P, C.c_int]),
    "fdfd_offdiag_fraction": (C.c_int, [P, C.POINTER(C.c_double)]),
    "fdfd_offdiag_bytes_per_dof": (C.c_int, [P, C.POINTER(C.c_double)]),
    "fdfd_offdiag_symmetric": (C.c_int, [P, C.POINTER(C.c_int)]),
    "fdfd_launch_count": (C.c_int64, [P]),
    "fdfd_host_alloc": (C.c_int, [C.POINTER(P), C.c_uint64]),
    "fdfd_host_free": (C.c_int, [P]),
    "fdfd_dev_alloc": (C.c_int, [P, C.POINTER(P), C.c_uint64]),
    "fdfd_dev_free": (C.c_int, [P, P]),
    "fdfd_memcpy": (C.c_int, [P, P, P, C.c_uint64, C.c_int, C.c_int]),
    # one call, N GPUs (multi.cpp)
    "fdfd_multi_create": (C.c_int, [C.POINTER(P), C.POINTER(Desc), C.c_int32, C.POINTER(C.c_int32)]),
    "fdfd_multi_destroy": (C.c_int, [P]),
    "fdfd_multi_last_error": (C.c_char_p, [P]),
    "fdfd_multi_ngpu": (C.c_int, [P]),
    "fdfd_multi_slab": (C.c_int, [P, C.c_int32, C.POINTER(P), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "fdfd_multi_set_coeffs": (C.c_int, [P, C.POINTER(P), C.POINTER(P)]),
    "fdfd_multi_set_bloch": (C.c_int, [P, P]),
    "fdfd_multi_set_omega": (C.c_int, [P, c128]),
    "fdfd_multi_set_eps": (C.c_int, [P, P, C.c_int]),
    "fdfd_multi_set_mu": (C.c_int, [P, P]),
    "fdfd_multi_set_eps_objects": (C.c_int, [P, P]),
    "fdfd_multi_apply": (C.c_int, [P, P, P]),
    "fdfd_multi_apply_transpose": (C.c_int, [P, P, P]),
    "fdfd_multi_solve": (C.c_int, [P, C.c_int, P, P, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double), P]),
    "fdfd_multi_create_b": (C.c_int, [P, P, P, P]),
    "fdfd_multi_h_from_e": (C.c_int, [P, P, P, P]),
    "fdfd_multi_e_from_h": (C.c_int, [P, P, P, P]),
}


def lib():
    """Load the CUDA shared library; raise loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                              f"g.build()'` (there is no CPU fallback)")
        l = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(l, name)
            f.restype, f.argtypes = res, args
        _lib = l
    return _lib


def check(code, handle=None, ok=(OK,)):
    if code not in ok:
        msg = lib().fdfd_last_error(handle)
        raise FdfdError(code, msg.decode() if msg else "")


def check_multi(code, handle=None, ok=(OK,)):
    if code not in ok:
        msg = lib().fdfd_multi_last_error(handle)
        raise FdfdError(code, msg.decode() if msg else "")
    return code
