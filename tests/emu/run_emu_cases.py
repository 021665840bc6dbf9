"""TEST INFRASTRUCTURE ONLY - runs groups of parity cases against build/emu/libfdfd_emu.so (the CUDA sources
compiled for the CPU, tests/emu/README.md) and checks every result against the oracle.

    FDFD_B200_LIB=build/emu/libfdfd_emu.so python tests/emu/run_emu_cases.py GROUP [GROUP ...] [--full]

Must run in its own process: the ctypes binding loads whatever FDFD_B200_LIB names, once.  "Device" buffers are host
memory here, so the FDFD_DEVICE entry points are called with numpy arrays.  Started by tests/test_emu_kernels_cpu.py
with FDFD_EMU_ASYNC / FDFD_EMU_SHUFFLE set to the scheduling mode under test."""
import ctypes as C
import itertools
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

from oracle.grid import EE, HH                      # noqa: E402
from oracle import operators as op                  # noqa: E402
from problems import Problem, rel                   # noqa: E402
import maxwellfdm_jl_b200 as fb                     # noqa: E402

L = fb._lib
TOL = 1e-12
NAIVE, TILED = 1, 2
FULL = "--full" in sys.argv


def apply_dev(A, x, transpose=False):
    x = np.ascontiguousarray(x, dtype=np.complex128)
    y = np.full(A.n, np.nan + 1j * np.nan)
    f = L.lib().fdfd_apply_transpose if transpose else L.lib().fdfd_apply
    L.check(f(A._h, x.ctypes.data, y.ctypes.data, L.DEVICE), A._h)
    return y


def check(err, what, tol=TOL):
    if not (err < tol):
        raise AssertionError(f"{what}: rel. error {err:.3e} (tol {tol:.0e})")


def g_apply():
    """every Bloch/symmetry combination, diagonal / full eps / +mu, odd shapes, N_w = 1..3 (tiled kernel vs CSC oracle)"""
    shapes = [(1, 1, 1), (2, 1, 3), (3, 2, 1), (3, 3, 2), (5, 3, 2), (3, 5, 8), (31, 15, 4), (33, 17, 9), (70, 45, 6)]
    combos = list(itertools.product([True, False], repeat=3))
    n = 0
    for N in shapes:
        for i, isbloch in enumerate(combos):
            for j, (full_eps, with_mu) in enumerate(((False, False), (True, False), (True, True))):
                if not FULL and (i + j + sum(N)) % 4 != 0:      # default: a quarter of the combinations per shape
                    continue
                p = Problem(N, isbloch, full_eps=full_eps, with_mu=with_mu)
                A_ref, _ = p.oracle_csc()
                A = p.operator(device=0, kernel=TILED)
                x = p.random_x()
                check(rel(apply_dev(A, x), A_ref.matvec(x)), f"apply {N} {isbloch} full={full_eps} mu={with_mu}")
                A.close()
                n += 1
    return n


def g_realmass():
    """real diagonal mass entries (lossless medium, real omega): material rows travel as doubles (MDR instantiations of
    the row-pair kernel) - every boundary combination, Bloch wrap tiles, several items per CTA, both uniform
    arrangements and a mixed one, full eps through the correction pass, transposed apply, fused dots via BiCGSTAB"""
    n = 0
    combos = list(itertools.product([True, False], repeat=3))
    for N in [(3, 3, 2), (33, 17, 9), (70, 45, 6), (31, 40, 5)]:
        for i, isbloch in enumerate(combos):
            if not FULL and (i + sum(N)) % 2 != 0:
                continue
            for full_eps in (False, True):
                p = Problem(N, isbloch, full_eps=full_eps, real_mass=True)
                A_ref, _ = p.oracle_csc()
                A = p.operator(device=0, kernel=TILED)
                x = p.random_x()
                check(rel(apply_dev(A, x), A_ref.matvec(x)), f"real-mass apply {N} {isbloch} full={full_eps}")
                check(rel(apply_dev(A, x, transpose=True), A_ref.to_scipy().T @ x), f"real-mass transpose {N} {isbloch}")
                A.close()
                n += 2
    for boundft in [(HH, HH, HH), (EE, HH, EE)]:
        for ft in (EE, HH):
            p = Problem((34, 19, 7), (True, False, True), boundft=boundft, ft=ft, with_mu=(ft == HH), real_mass=True)
            A_ref, _ = p.oracle_csc()
            A = p.operator(device=0, kernel=TILED)
            x = p.random_x()
            check(rel(apply_dev(A, x), A_ref.matvec(x)), f"real-mass apply boundft={boundft} ft={ft}")
            A.close()
            n += 1
    p = Problem((12, 9, 6), (True, True, False), real_mass=True, npml=2)
    A_ref, _ = p.oracle_csc()
    A = p.operator(device=0, kernel=TILED)
    b = A_ref.matvec(p.random_x(3))
    xs = np.zeros(A.n, complex)
    iters, relres = C.c_int(), C.c_double()
    L.check(L.lib().fdfd_solve(A._h, L.BICGSTAB, b.ctypes.data, xs.ctypes.data, L.DEVICE, 1e-10, 4000, 5,
                               C.byref(iters), C.byref(relres), None), A._h, ok=(L.OK, L.ENOCONV))
    check(rel(A_ref.matvec(xs), b), "real-mass BiCGSTAB true residual", 1e-8)
    A.close()
    return n + 1


def g_fused():
    """fused full-tensor shape of the row-pair kernel (real diagonal + real symmetric off-diagonal entries through the TMA
    ring): every boundary combination incl. Bloch wrap at the forward tile edge, dense / partly empty / empty
    off-diagonal blocks (tile mask and per-row-pair skip), several z-chunks, uniform and mixed arrangements, both
    formulations, transposed apply, fused dots through BiCGSTAB"""
    n = 0
    combos = list(itertools.product([True, False], repeat=3))
    os.environ.pop("FDFD_RP_FUSE_MIN", None)
    for N in [(3, 3, 2), (33, 17, 9), (70, 45, 6), (31, 40, 5), (61, 29, 7)]:
        for i, isbloch in enumerate(combos):
            if not FULL and (i + sum(N)) % 2 != 0:
                continue
            p = Problem(N, isbloch, full_eps=True, real_mass=True, sym_real_off=True)
            A_ref, _ = p.oracle_csc()
            A = p.operator(device=0, kernel=TILED)
            x = p.random_x()
            check(rel(apply_dev(A, x), A_ref.matvec(x)), f"fused apply {N} {isbloch}")
            check(rel(apply_dev(A, x, transpose=True), A_ref.to_scipy().T @ x), f"fused transpose {N} {isbloch}")
            A.close()
            n += 2
    # partly empty off-diagonal blocks: whole planes, a half-space in y (row pairs skip), a half-space in x
    cases = [((EE, EE, EE), (True, True, True), EE), ((HH, HH, HH), (False, True, False), EE), ((EE, HH, EE), (True, False, True), EE),
             ((EE, EE, EE), (False, False, True), HH)]
    for boundft, isbloch, ft in (cases if FULL else cases[:3]):
        p = Problem((64, 45, 40), isbloch, boundft, full_eps=(ft == EE), full_mu=(ft == HH), with_mu=(ft == HH), ft=ft,
                    real_mass=True, sym_real_off=True)
        mass = p.eps if ft == EE else p.mu
        for v, u in itertools.permutations(range(3), 2):
            mass[:, :, :9, v, u] = 0
            mass[:, :, 21:30, v, u] = 0
            mass[:, 11:30, 30:, v, u] = 0
            mass[:33, :, 12:18, v, u] = 0
        x = p.random_x()
        A = p.operator(device=0, kernel=TILED)
        An = p.operator(device=0, kernel=NAIVE)
        y = apply_dev(A, x)
        check(rel(y, apply_dev(An, x)), f"fused vs general kernel {boundft} {isbloch} ft={ft}", 1e-13)
        check(rel(y, p.oracle_matfree()(x)), f"fused vs matrix-free oracle {boundft} {isbloch} ft={ft}")
        A.close()
        An.close()
        n += 1
    p = Problem((12, 9, 6), (True, True, False), full_eps=True, real_mass=True, sym_real_off=True, npml=2)
    A_ref, _ = p.oracle_csc()
    A = p.operator(device=0, kernel=TILED)
    b = A_ref.matvec(p.random_x(3))
    xs = np.zeros(A.n, complex)
    iters, relres = C.c_int(), C.c_double()
    L.check(L.lib().fdfd_solve(A._h, L.BICGSTAB, b.ctypes.data, xs.ctypes.data, L.DEVICE, 1e-10, 4000, 5,
                               C.byref(iters), C.byref(relres), None), A._h, ok=(L.OK, L.ENOCONV))
    check(rel(A_ref.matvec(xs), b), "fused BiCGSTAB true residual", 1e-8)
    A.close()
    return n + 1


def g_multi():
    """the single-call handle (fdfd_multi_*, csrc/multi.cpp) with ONE slab - what can run without several devices: full-grid
    host arrays in, apply / transposed apply on both DOF layouts, BiCGSTAB / QMR, the model-level call create_A(..., ngpu=1);
    several slabs (host threads, NCCL between them) are checked on the GPU boxes by scripts/multi_check.py"""
    import types
    sys.modules.setdefault("torch", types.SimpleNamespace(cuda=types.SimpleNamespace(device_count=lambda: 1)))
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import multi_check
    n, fails = multi_check.check(1)
    if fails:
        raise AssertionError(f"single-call handle: {fails[:3]}")
    return n


def g_boundft():
    """all 2^3 boundft choices, both formulations, forward and transposed (ARR = 0 / 1 / 2 instantiations)"""
    n = 0
    for boundft in itertools.product([EE, HH], repeat=3):
        for isbloch in ((True, False, True), (False, True, False)):
            for ft in (EE, HH):
                p = Problem((9, 6, 7), isbloch, boundft, full_eps=(ft == EE), with_mu=True, ft=ft, full_mu=(ft == HH))
                A_ref, _ = p.oracle_csc()
                A = p.operator(device=0, kernel=TILED)
                x = p.random_x()
                check(rel(apply_dev(A, x), A_ref.matvec(x)), f"boundft {boundft} {isbloch} ft={ft}")
                check(rel(apply_dev(A, x, True), A_ref.to_scipy().T.tocsc() @ x), f"boundft^T {boundft} {isbloch} ft={ft}")
                A.close()
                n += 2
    return n


def g_layouts():
    """component-major layout, weighted output average, omega = 0, host path == device path"""
    n = 0
    for cmpfirst, wo in ((False, False), (True, True), (False, True)):
        p = Problem((37, 20, 11), (True, False, True), full_eps=True, with_mu=True, cmpfirst=cmpfirst, weighted_out=wo)
        A_ref, _ = p.oracle_csc()
        for k in (TILED, NAIVE):
            A = p.operator(device=0, kernel=k)
            x = p.random_x()
            check(rel(apply_dev(A, x), A_ref.matvec(x)), f"layout cmpfirst={cmpfirst} wo={wo} kernel={k}")
            A.close()
            n += 1
    for sym in (True, False):                # pointwise symmetric off-diagonals are stored once (aliased slots)
        p = Problem((37, 20, 11), (True, False, True), full_eps=True, with_mu=True)
        if sym:
            for v, u in itertools.combinations(range(3), 2):
                p.eps[..., u, v] = p.eps[..., v, u]
        A_ref, _ = p.oracle_csc()
        for k in (TILED, NAIVE):
            A = p.operator(device=0, kernel=k)
            assert A.offdiag_symmetric == sym
            x = p.random_x()
            check(rel(apply_dev(A, x), A_ref.matvec(x)), f"symmetric={sym} kernel={k}")
            check(rel(apply_dev(A, x, True), A_ref.to_scipy().T.tocsc() @ x), f"symmetric={sym} kernel={k} transposed")
            A.close()
            n += 2
    p = Problem((10, 9, 8), (True, False, True), omega=0.0)
    A_ref, _ = p.oracle_csc()
    A = p.operator(device=0, kernel=TILED)
    x = p.random_x()
    check(rel(apply_dev(A, x), A_ref.matvec(x)), "omega = 0")
    A.close()
    p = Problem((20, 18, 10), (False, True, False), full_eps=True)
    A = p.operator(device=0)
    x = p.random_x()
    assert np.array_equal(A @ x, apply_dev(A, x)), "host-buffer path differs from the device path"
    A.close()
    return n + 2


def g_deep():
    """several tiles and z-chunks per column, sparse / dense / absent off-diagonal blocks (occupancy mask, marching
    correction kernel, fused kernel), uniform and mixed arrangements, both layouts; vs the matrix-free oracle and
    the general kernel"""
    n = 0
    cases = [((EE, EE, EE), (True, False, True), 33, True), ((EE, EE, EE), (False, False, False), 60, True),
             ((HH, HH, HH), (False, True, False), 33, True), ((HH, HH, HH), (True, True, True), 60, False),
             ((EE, HH, EE), (True, False, True), 33, True), ((HH, EE, HH), (False, True, False), 60, False)]
    if not FULL:
        cases = cases[:1] + cases[3:5]
    for boundft, isbloch, z1, cmpfirst in cases:
        p = Problem((70, 45, 90), isbloch, boundft, full_eps=True, with_mu=True, cmpfirst=cmpfirst)
        for v, u in itertools.permutations(range(3), 2):
            p.eps[:, :, :20, v, u] = 0
            p.eps[:, :, z1:, v, u] = 0
            p.eps[:25, :, :, v, u] = 0
        x = p.random_x()
        At = p.operator(device=0, kernel=TILED)
        An = p.operator(device=0, kernel=NAIVE)
        yt, yn = apply_dev(At, x), apply_dev(An, x)
        check(rel(yt, yn), f"deep tiled vs general {boundft} {isbloch}", 1e-13)
        check(rel(yt, p.oracle_matfree()(x)), f"deep vs matrix-free oracle {boundft} {isbloch}")
        At.close()
        An.close()
        n += 1
    p = Problem((40, 33, 70), (False, True, False), ft=HH, with_mu=True)
    A = p.operator(device=0, kernel=TILED)
    x = p.random_x()
    check(rel(apply_dev(A, x), p.oracle_matfree()(x)), "deep HH formulation")
    A.close()
    return n + 1


def g_solve():
    """BiCGSTAB (with the fused apply-epilogue dots) and QMR on the emulated device vs a sparse direct solve"""
    import scipy.sparse.linalg as spla
    n = 0
    for ft, kw in ((EE, dict(full_eps=True)), (HH, dict(with_mu=True))):
        p = Problem((12, 10, 9), (True, False, True), ft=ft, omega=1.3 - 0.4j, **kw)
        A_ref, _ = p.oracle_csc()
        b = A_ref.matvec(p.random_x(5))
        x_ref = spla.splu(A_ref.to_scipy().tocsc()).solve(b)
        A = p.operator(device=0, kernel=TILED)
        for method in ("bicgstab", "qmr"):
            x = np.zeros(A.n, complex)
            import ctypes as C
            iters, relres = C.c_int(), C.c_double()
            m = L.BICGSTAB if method == "bicgstab" else L.QMR
            code = L.lib().fdfd_solve(A._h, m, b.ctypes.data, x.ctypes.data, L.DEVICE, 1e-10, 400, 10, C.byref(iters),
                                      C.byref(relres), None)
            L.check(code, A._h)
            check(rel(A_ref.matvec(x), b), f"{method} true residual ft={ft}", 1e-8)
            check(rel(x, x_ref), f"{method} field vs direct solve ft={ft}", 1e-7)
            n += 1
        A.close()
    return n


def g_aux():
    """create_b, h_from_e, e_from_h, corner interpolation on EE and HH handles vs the oracle"""
    n = 0
    for isbloch, boundft, ft in (((True, False, True), (EE, EE, EE), EE), ((True, True, False), (HH, EE, HH), EE),
                                 ((False, True, False), (HH, EE, HH), HH)):
        p = Problem((9, 8, 7), isbloch, boundft, with_mu=True, ft=ft)
        A_ref, (Pe, Pm, Ce, Cm) = p.oracle_csc()
        A = p.operator(device=0)
        je, jm, e, h = p.random_x(31), p.random_x(32), p.random_x(33), p.random_x(34)
        check(rel(A.create_b(je, jm), op.create_b(ft, p.omega, Pe, Pm, Ce, Cm, je, jm)), "create_b")
        check(rel(A.e_from_h(h, je), op.e_from_h(h, p.omega, Pe, Cm, je)), "e_from_h")
        check(rel(A.h_from_e(e, jm), op.h_from_e(e, p.omega, Pm, Ce, jm)), "h_from_e")
        Mce, Mcm = op.create_Mcs(p.sdl_e, p.sdl_m, p.sei, p.smi, boundft, isbloch, p.ph)
        check(rel(A.interp_corners(e, "E"), Mce.matvec(e)), "interp_corners E")
        check(rel(A.interp_corners(h, "H"), Mcm.matvec(h)), "interp_corners H")
        A.close()
        n += 5
    return n


def g_matparams():
    """N4: object assignment + Kottke smoothing kernel vs the oracle (boxes, balls, cylinders; Bloch / symmetry ghost
    corners; non-uniform grids; full-tensor materials; mu locations; z-slabs)"""
    from oracle import matparams as omp
    from oracle.grid import Grid as OGrid
    from problems import matparams_scene, MATPARAMS_CASES
    n = 0
    for N, isbloch, boundft, ft, uniform, nshape, aniso in MATPARAMS_CASES:
        lp, o_sh, f_sh, pinds, params = matparams_scene(N, isbloch, uniform, nshape, aniso)
        ref = omp.calc_matparams(OGrid(lp, isbloch), boundft, ft, o_sh, pinds, params)
        g = fb.Grid(lp, isbloch)
        got = fb.calc_matparams_array(g, boundft, ft, f_sh, pinds, params, device=0)
        check(rel(got, ref), f"matparams {N} {isbloch} {boundft} ft={ft}")
        k0, k1 = 2, N[2] - 1                                       # a z-slab is the same planes of the same array
        slab = fb.calc_matparams_array(g, boundft, ft, f_sh, pinds, params, k0=k0, k1=k1, device=0)
        assert np.array_equal(slab, got[:, :, k0:k1]), "slab differs from the full-grid result"
        n += 2
    # eps straight from objects (no host array): same operator as the one fed with the array calc_matparams_array returns
    for N, isbloch, boundft_o, sym in (((12, 10, 9), (True, False, True), (EE, EE, EE), True), ((13, 9, 11), (False, True, False), (EE, HH, EE), False)):
        lp, o_sh, f_sh2, pinds2, params2 = matparams_scene(N, isbloch, False, 6, not sym)
        p = Problem(N, isbloch, boundft_o)
        p.grid_lp = lp
        g2 = fb.Grid(lp, isbloch)
        eps = fb.calc_matparams_array(g2, boundft_o, EE, f_sh2, pinds2, params2, device=0)
        bf = ["E" if b == EE else "H" for b in boundft_o]
        A1 = fb.FdfdOperator(N, isbloch, p.sdl_e, p.sdl_m, p.omega, eps, None, p.ph, boundft=bf, device=0, kernel=TILED)
        A2 = fb.FdfdOperator(N, isbloch, p.sdl_e, p.sdl_m, p.omega, None, None, p.ph, boundft=bf, device=0, kernel=TILED)
        A2.set_eps_objects(lp, f_sh2, pinds2, params2, boundft=bf)
        x = p.random_x()
        y1, y2 = apply_dev(A1, x), apply_dev(A2, x)
        assert np.array_equal(y1, y2), f"objects path differs from the array path: {rel(y2, y1):.2e}"
        assert A1.offdiag_symmetric == sym and A2.offdiag_symmetric == sym, (A1.offdiag_symmetric, A2.offdiag_symmetric, sym)
        assert np.isfinite(y1).all() and A2.offdiag_fraction > 0
        try:
            A2.export_pattern()
        except L.FdfdError as e:
            assert e.code == L.ESTATE
        else:
            raise AssertionError("export_pattern must refuse a handle without a host eps array")
        A1.close()
        A2.close()
        n += 2
    try:
        fb.calc_matparams_array(g, boundft, ft, f_sh[1:2], [0], params[:1], device=0)
    except L.FdfdError as e:
        assert e.code == L.EINVAL and "covered by no shape" in str(e)
        n += 1
    else:
        raise AssertionError("an uncovered grid must be rejected")
    return n


def g_reduced():
    """ModelTE / ModelTM / ModelTEM (maxwellfdm.jl_b200/reduced.py: 3-D handle one periodic cell thick) against the
    K-dimensional oracle"""
    from problems import REDUCED_CASES, reduced_model_check, reduced_objects_check
    n = reduced_objects_check(fb)
    # the torch branch of ReducedOperator (embedding / extraction with tensor indexing) on CPU tensors: the emulation
    # build takes host pointers, the GPU box runs the same code on CUDA tensors (test_reduced_operator_on_device_tensors)
    import torch
    rng = np.random.default_rng(1)
    for cmpfirst in (True, False):
        mdl = fb.ModelTE(fb.Grid([np.arange(8.0), np.arange(6.0)], (True, False)))
        mdl.order_cmpfirst = cmpfirst
        mdl.eps_arr[..., 0, 0], mdl.eps_arr[..., 1, 1] = 2.0, 3.0
        mdl.eps_arr[..., 0, 1] = mdl.eps_arr[..., 1, 0] = 0.1
        A = fb.create_A(0, 1.0, mdl, device=0)
        x = rng.standard_normal(A.n) + 1j * rng.standard_normal(A.n)
        y = A @ x
        yt = A @ torch.from_numpy(x)
        assert isinstance(yt, torch.Tensor) and rel(yt.numpy(), y) == 0.0
        out = torch.empty(A.n, dtype=torch.complex128)
        A.mul(out, torch.from_numpy(x))
        assert rel(out.numpy(), y) == 0.0
        xs, info = A.solve(torch.from_numpy(y), rtol=1e-12)
        assert info["converged"] and rel(xs.numpy(), x) < 1e-9
        A.close()
        n += 3
    for case in REDUCED_CASES:
        errs = reduced_model_check(fb, *case)
        assert max(v for k, v in errs.items() if k != "solve") < 1e-12 and errs["solve"] < 1e-7, (case, errs)
        n += len(errs)
    return n


GROUPS = {"multi": g_multi, "fused": g_fused, "realmass": g_realmass, "reduced": g_reduced, "matparams": g_matparams, "apply": g_apply, "boundft": g_boundft, "layouts": g_layouts, "deep": g_deep, "solve": g_solve, "aux": g_aux}


def main():
    ver = L.lib().fdfd_version().decode()
    if "EMULATED" not in ver:
        sys.exit("run_emu_cases.py must be pointed at the emulation build (FDFD_B200_LIB=build/emu/libfdfd_emu.so); "
                 f"loaded: {ver}")
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or list(GROUPS)
    for name in names:
        t = time.time()
        n = GROUPS[name]()
        print(f"emu[{os.environ.get('FDFD_EMU_ASYNC', 'eager')},shuffle={os.environ.get('FDFD_EMU_SHUFFLE', '0')}] "
              f"{name}: {n} checks ok in {time.time() - t:.1f}s", flush=True)


if __name__ == "__main__":
    main()
