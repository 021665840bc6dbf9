"""CPU-side checks of the C-ABI library (no compute calls): it loads, exports every symbol the header
declares, refuses to run without a GPU, and its host-only debug export reproduces the oracle's sparse
index pattern BIT-EXACTLY (reference-defined integer work) and its values to 1e-13."""
import ctypes as C
import itertools
import os
import re

import numpy as np
import pytest

from oracle.grid import EE, HH
from problems import Problem, rel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    import maxwellfdm_jl_b200 as fb
    return fb._lib


def test_library_exports_every_declared_symbol():
    L = _lib()
    hdr = open(os.path.join(ROOT, "include", "fdfd_b200.h")).read()
    declared = set(re.findall(r"\b(fdfd_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    lib = L.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/fdfd_b200.h but not exported"
    assert declared == set(L.SYMBOLS), declared ^ set(L.SYMBOLS)
    assert b"sm_100a" in lib.fdfd_version()


def test_plain_c_consumer_and_struct_layouts(tmp_path):
    """tests/cabi_smoke.c: the header compiles as plain C, the library links from C, and the struct layouts a foreign
    binding must reproduce (ctypes here, the Julia struct in julia/FDFDB200.jl) are the C compiler's."""
    import subprocess
    L = _lib()
    exe = str(tmp_path / "cabi_smoke")
    libdir = os.path.dirname(L.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cabi_smoke.c"),
                    "-o", exe, "-L", libdir, "-l:" + os.path.basename(L.LIB_PATH), "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert "cabi_smoke ok" in out, out
    offs = dict(l.split() for l in out.splitlines() if re.match(r"^(fdfd_\w+\.\w+|sizeof\.\w+) \d+$", l))
    for cname, T in (("fdfd_desc", L.Desc), ("fdfd_shape", L.Shape), ("fdfd_matparams_desc", L.MatParamsDesc), ("fdfd_c128", L.c128)):
        assert int(offs["sizeof." + cname]) == C.sizeof(T), cname
        for f, _ in T._fields_:
            key = f"{cname}.{f}"
            if key in offs:
                assert int(offs[key]) == getattr(T, f).offset, key
    assert sum(k.startswith("fdfd_desc.") for k in offs) == len(L.Desc._fields_)
    # the Julia binding lists the same fields in the same order (isbits struct = C layout)
    jl = open(os.path.join(ROOT, "julia", "FDFDB200.jl")).read()
    body = re.search(r"\nstruct Desc\n(.*?)\nend", jl, re.S).group(1)
    jl_fields = re.findall(r"^\s*(\w+)::", body, re.M)
    assert jl_fields == [f for f, _ in L.Desc._fields_], jl_fields


def test_partition_rule():
    import maxwellfdm_jl_b200 as fb
    for Nz, P in [(1, 1), (7, 3), (768, 8), (512, 8), (5, 5), (200, 7)]:
        edges = [fb.partition(Nz, P, r) for r in range(P)]
        assert edges[0][0] == 0 and edges[-1][1] == Nz
        for a, b in zip(edges[:-1], edges[1:]):
            assert a[1] == b[0]
        sizes = [b - a for a, b in edges]
        assert min(sizes) >= 1 and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        fb.partition(3, 4, 0)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = _lib()
    p = Problem((3, 3, 3))
    with pytest.raises(L.FdfdError) as ei:
        p.operator()
    assert ei.value.code == L.ECUDA and "no CPU fallback" in str(ei.value)
    # the material pipeline is a GPU kernel too: no device, no result
    import maxwellfdm_jl_b200 as fb
    lp = np.arange(5.0)
    with pytest.raises(L.FdfdError) as ei2:
        fb.calc_matparams_array(fb.Grid((lp, lp, lp), (True,) * 3), (EE,) * 3, EE, [fb.Box([2, 2, 2], [4, 4, 4])], [0],
                                [np.eye(3)])
    assert ei2.value.code == L.ECUDA
    # a host-only handle exports patterns but refuses every compute call
    A = p.operator(device=-2)
    with pytest.raises(L.FdfdError) as ei:
        A @ p.random_x()
    assert ei.value.code == L.ESTATE
    with pytest.raises(L.FdfdError):
        A.solve(p.random_x())


def test_single_call_handle_needs_gpus_too():
    """fdfd_multi_* owns ordinary slab handles: without a CUDA device it must fail loudly (no CPU fallback), with the failing
    slab named in the message"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import maxwellfdm_jl_b200 as fb
    L = _lib()
    p = Problem((4, 3, 6))
    with pytest.raises(L.FdfdError) as ei:
        fb.MultiGpuOperator(p.N, p.isbloch, p.sdl_e, p.sdl_m, p.omega, p.eps, None, p.ph, ngpu=2)
    assert ei.value.code == L.ECUDA and "no CPU fallback" in str(ei.value) and "slab" in str(ei.value)
    with pytest.raises(L.FdfdError) as ei2:        # more slabs than z-planes
        fb.MultiGpuOperator(p.N, p.isbloch, p.sdl_e, p.sdl_m, p.omega, p.eps, None, p.ph, ngpu=7)
    assert ei2.value.code == L.EINVAL


def test_output_buffers_are_validated_before_the_c_call():
    """mul!(y, A, x) / solve(b, x0): the library writes y / reads x0 through raw pointers, so the Python mirror rejects
    a wrong dtype, size, stride or a host / device mix instead of handing it to the C ABI."""
    import torch
    L = _lib()
    p = Problem((4, 3, 2), (True, True, True))
    A = p.operator(device=-2)          # validation happens before the C call, so no GPU is needed to see it
    x = p.random_x()
    for bad in (np.empty(A.n, np.float64), np.empty(A.n - 1, np.complex128), np.empty(2 * A.n, np.complex128)[::2],
                [0j] * A.n, torch.empty(A.n, dtype=torch.complex128)):
        with pytest.raises(ValueError):
            A.mul(bad, x)
    with pytest.raises(ValueError):
        A.mul(np.empty(A.n, np.complex128), torch.from_numpy(x))      # torch input, numpy output
    with pytest.raises(ValueError):
        A.solve(x, x0=torch.zeros(A.n, dtype=torch.complex128))       # numpy b, torch x0
    with pytest.raises(L.FdfdError):                                  # a well-formed call reaches the library
        A.mul(np.empty(A.n, np.complex128), x)


def _check_export(p):
    A, _ = p.oracle_csc()
    cp_ref, rv_ref = A.julia_pattern()
    op_ = p.operator(device=-2)
    cp, rv, nz = op_.export_pattern()
    assert cp.dtype == np.int64 and rv.dtype == np.int64
    assert np.array_equal(cp, cp_ref), "colptr differs"
    assert np.array_equal(rv, rv_ref), "rowval differs"
    scale = np.abs(A.nzval).max()
    assert np.abs(nz - A.nzval).max() <= 1e-13 * scale
    op_.close()


SIZES = [(1, 1, 1), (2, 1, 3), (3, 2, 1), (5, 3, 2), (3, 5, 8)]


@pytest.mark.parametrize("N", SIZES)
@pytest.mark.parametrize("isbloch", list(itertools.product([True, False], repeat=3)))
def test_export_pattern_default_boundft(N, isbloch):
    for full_eps, with_mu in ((False, False), (True, True)):
        _check_export(Problem(N, isbloch, full_eps=full_eps, with_mu=with_mu))


@pytest.mark.parametrize("boundft", list(itertools.product([EE, HH], repeat=3)))
def test_export_pattern_all_boundft(boundft):
    for isbloch in ((True, False, True), (False, True, False)):
        for full_eps in (False, True):
            _check_export(Problem((4, 3, 5), isbloch, boundft, full_eps=full_eps, with_mu=True))


def test_export_pattern_variants():
    # w == 0: structural pattern with explicit zeros kept (model.jl:237 skips the subtraction)
    for isbloch in ((True,) * 3, (False,) * 3):
        p = Problem((4, 5, 3), isbloch, omega=0.0, uniform=True, npml=0)
        A, _ = p.oracle_csc()
        cp, rv, nz = p.operator(device=-2).export_pattern()
        assert np.array_equal(cp, A.julia_pattern()[0]) and np.array_equal(rv, A.julia_pattern()[1])
        assert np.all(np.diff(cp) == 13)
        if not any(isbloch):
            assert (nz == 0).any()
    # component-major DOF order, weighted output average, HH formulation
    _check_export(Problem((4, 3, 5), (True, False, True), full_eps=True, with_mu=True, cmpfirst=False))
    _check_export(Problem((4, 3, 5), (False, True, True), full_eps=True, weighted_out=True))
    _check_export(Problem((4, 3, 5), (True, False, True), with_mu=True, ft=HH))
    _check_export(Problem((4, 3, 5), (True, True, False), (HH, EE, HH), with_mu=True, ft=HH, cmpfirst=False))
    # HH formulation with a full 3x3 mu tensor as the mass parameter (model.jl:238-240)
    _check_export(Problem((4, 3, 5), (False, True, True), ft=HH, full_mu=True))
    _check_export(Problem((3, 4, 2), (True, False, False), (HH, HH, EE), ft=HH, full_mu=True, weighted_out=True))
    # exact zeros are dropped when w != 0 (symmetry boundaries on a uniform grid)
    p = Problem((5, 6, 7), (False,) * 3, uniform=True, npml=0, omega=1.0)
    A, _ = p.oracle_csc()
    cp, rv, nz = p.operator(device=-2).export_pattern()
    assert np.array_equal(rv, A.julia_pattern()[1]) and np.all(nz != 0) and (np.diff(cp) < 13).any()


@pytest.mark.parametrize("cfg", ["C1 40^3", "C2 reduced", "C3 reduced"])
def test_export_pattern_at_config_shapes(cfg):
    """the bit-exact index pattern on the BASELINE configurations themselves - C1 at its full 40^3 size
    (192 000 unknowns, 10-cell PML), C2 (full 3x3 eps, PML) and C3 (Bloch x / y with a complex phase, PML in z) at
    reduced sizes - built by the same workload generators the bench uses (model.jl:236-237 is what the export mirrors)."""
    import sys
    sys.path.insert(0, ROOT)
    import workloads
    from oracle import operators as op
    w = {"C1 40^3": lambda: workloads.c1_vacuum_box((40, 40, 40)),
         "C2 reduced": lambda: workloads.c2_waveguide((36, 40, 30), npml=6),
         "C3 reduced": lambda: workloads.c3_phc_slab((32, 32, 24))}[cfg]()
    sei = tuple(1 / a for a in w["sdl_e"])
    smi = tuple(1 / a for a in w["sdl_m"])
    mu = np.zeros(w["eps"].shape, complex)
    for v in range(3):
        mu[..., v, v] = 1
    Ce, Cm = op.create_curls(sei, smi, (EE, EE, EE), w["isbloch"], w["e_mikL"])
    Pe, Pm = op.create_paramops(w["eps"], mu, w["sdl_e"], w["sdl_m"], sei, smi, (EE, EE, EE), w["isbloch"], w["e_mikL"])
    A = op.create_A(EE, w["omega"], Pe, Pm, Ce, Cm)
    cp_ref, rv_ref = A.julia_pattern()
    H = workloads.make_operator(w, device=-2)
    cp, rv, nz = H.export_pattern()
    H.close()
    assert np.array_equal(cp, cp_ref) and np.array_equal(rv, rv_ref), cfg
    assert np.abs(nz - A.nzval).max() <= 1e-13 * np.abs(A.nzval).max()


def test_export_capacity_and_errors():
    L = _lib()
    p = Problem((3, 3, 3))
    A = p.operator(device=-2)
    nnz = C.c_int64(5)
    cp = np.zeros(A.n + 1, np.int64)
    rv = np.zeros(5, np.int64)
    code = L.lib().fdfd_export_pattern(A._h, cp.ctypes.data, rv.ctypes.data, None, C.byref(nnz))
    assert code == L.EINVAL and nnz.value == 13 * A.n
    d = L.Desc()
    d.N[:] = [3, 3, 0]
    h = C.c_void_p()
    assert L.lib().fdfd_create(C.byref(h), C.byref(d)) == L.EINVAL
    with pytest.raises(L.FdfdError):       # mu with off-diagonal entries (reference model.jl:236)
        mu = p.mu.copy()
        mu[..., 0, 1] = 0.1
        A.set_mu(mu)


def test_export_pattern_of_reduced_models():
    """ModelTE / ModelTM / ModelTEM: the index pattern of A (1-based Int64, Julia CSC order) is the block of the 3-D
    export and equals the K-dimensional operator's structure bit for bit (host-only handle: no GPU needed)."""
    import maxwellfdm_jl_b200 as fb
    from problems import REDUCED_PATTERN_CASES, reduced_pattern_check
    for case in REDUCED_PATTERN_CASES:
        assert reduced_pattern_check(fb, *case) < 1e-14, case
