"""ctypes wrapper of oracle/csc_spmv.c (CPU baseline: Julia-style CSC mul! and an OpenMP CSR variant)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libcscspmv.so")
_lib = None


def build():
    subprocess.run(["make", "-C", HERE, "-s"], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            build()
        l = C.CDLL(SO)
        P = C.c_void_p
        l.csc_mul_serial.argtypes = [C.c_int64, P, P, P, P, P]
        l.csr_mul_omp.argtypes = [C.c_int64, P, P, P, P, P]
        l.bicgstab_csr_omp.argtypes = [C.c_int64, P, P, P, P, P, C.c_int, P]
        l.bicgstab_csr_omp.restype = C.c_double
        l.oracle_num_threads.restype = C.c_int
        l.oracle_set_threads.argtypes = [C.c_int]
        _lib = l
    return _lib


def csc_mul_serial(A, x, y=None):
    """y = A x with A an oracle.operators.Csc (0-based int64 indices)."""
    n = A.shape[0]
    y = np.empty(n, np.complex128) if y is None else y
    cp = np.ascontiguousarray(A.colptr, np.int64)
    rv = np.ascontiguousarray(A.rowval, np.int64)
    nz = np.ascontiguousarray(A.nzval, np.complex128)
    x = np.ascontiguousarray(x, np.complex128)
    lib().csc_mul_serial(n, cp.ctypes.data, rv.ctypes.data, nz.ctypes.data, x.ctypes.data, y.ctypes.data)
    return y


class CsrOmp:
    """Row-major copy of a Csc for the all-cores product."""

    def __init__(self, A):
        S = A.to_scipy().tocsr()
        S.sort_indices()
        self.n = A.shape[0]
        self.rowptr = np.ascontiguousarray(S.indptr, np.int64)
        self.colval = np.ascontiguousarray(S.indices, np.int64)
        self.nzval = np.ascontiguousarray(S.data, np.complex128)

    def mul(self, x, y):
        lib().csr_mul_omp(self.n, self.rowptr.ctypes.data, self.colval.ctypes.data, self.nzval.ctypes.data,
                          x.ctypes.data, y.ctypes.data)
        return y


def _bicgstab(self, b, x, iters):
    """x <- BiCGSTAB iterate after exactly `iters` iterations from x (no convergence exit); returns ||r|| / ||b||"""
    b = np.ascontiguousarray(b, np.complex128)
    work = np.empty(6 * self.n, np.complex128)
    return float(lib().bicgstab_csr_omp(self.n, self.rowptr.ctypes.data, self.colval.ctypes.data, self.nzval.ctypes.data,
                                        b.ctypes.data, x.ctypes.data, int(iters), work.ctypes.data))


CsrOmp.bicgstab = _bicgstab


def num_threads():
    return int(lib().oracle_num_threads())


def use_all_cores():
    """Size the OpenMP team to the cores this process may run on, whatever OMP_NUM_THREADS says (torchrun sets
    it to 1 for its workers).  Returns the team size."""
    import os
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().oracle_set_threads(int(n))
    return num_threads()
