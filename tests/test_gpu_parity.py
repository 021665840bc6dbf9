"""GPU parity tests proper (run with -m gpu on the B200 box): the CUDA path, called through the C ABI,
against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): A*x within 1e-12 relative of the oracle's CSC product;
index pattern bit-exact; converged fields within 1e-8 relative.
"""
import itertools
import os

import numpy as np
import pytest

from oracle.grid import EE, HH
from oracle import operators as op
from problems import Problem, rel, crandn, SEED

pytestmark = pytest.mark.gpu

TOL = 1e-12


def _fb():
    import maxwellfdm_jl_b200 as fb
    return fb


def _torch():
    import torch
    return torch


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    torch = _torch()
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)


def _apply_dev(A, x, transpose=False):
    torch = _torch()
    xd = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    yd = torch.empty_like(xd)
    A.mul(yd, xd, transpose=transpose)
    return yd.cpu().numpy()


KERNELS = {"naive": 1, "tiled": 2}

SMALL = [(1, 1, 1), (2, 1, 3), (3, 2, 1), (3, 3, 2), (5, 3, 2), (3, 5, 8), (31, 15, 4), (33, 17, 9), (70, 45, 6)]


@pytest.mark.parametrize("kernel", ["tiled", "naive"])
@pytest.mark.parametrize("N", SMALL)
def test_apply_all_boundary_conditions(kernel, N):
    """every Bloch/symmetry combination, diagonal and full eps, with and without mu, odd shapes that are
    not multiples of the tile, N_w = 1..3 edge cases (SURVEY §8c parity procedure (ii))."""
    for isbloch in itertools.product([True, False], repeat=3):
        for full_eps, with_mu in ((False, False), (True, False), (True, True)):
            p = Problem(N, isbloch, full_eps=full_eps, with_mu=with_mu)
            A_ref, _ = p.oracle_csc()
            A = p.operator(device=0, kernel=KERNELS[kernel])
            x = p.random_x()
            err = rel(_apply_dev(A, x), A_ref.matvec(x))
            A.close()
            assert err < TOL, (kernel, N, isbloch, full_eps, with_mu, err)


@pytest.mark.parametrize("N", [(3, 3, 2), (33, 17, 9), (70, 45, 6), (31, 40, 5), (64, 50, 40)])
def test_apply_real_mass_rows(N):
    """lossless medium at a real frequency: the diagonal mass entries are real and the row-pair kernel streams them as
    doubles (MDR instantiations; odd Nx exercises the padded row pitch of the tensor map) - every boundary combination,
    diagonal and full eps (correction pass on top), forward and transposed, against the oracle's CSC product."""
    for isbloch in itertools.product([True, False], repeat=3):
        for full_eps in (False, True):
            p = Problem(N, isbloch, full_eps=full_eps, real_mass=True)
            A_ref, _ = p.oracle_csc()
            A = p.operator(device=0, kernel=KERNELS["tiled"])
            x = p.random_x()
            err = rel(_apply_dev(A, x), A_ref.matvec(x))
            errT = rel(_apply_dev(A, x, transpose=True), A_ref.to_scipy().T.tocsc() @ x)
            A.close()
            assert err < TOL and errT < TOL, (N, isbloch, full_eps, err, errT)


def test_real_mass_rows_other_arrangements_and_solve():
    """MDR on the mirrored and a mixed arrangement, both formulations; BiCGSTAB (fused dots in the apply epilogue) and
    QMR (transposed operator) on a real-mass problem against a sparse direct solve."""
    import scipy.sparse.linalg as spla
    torch = _torch()
    for boundft in [(HH, HH, HH), (EE, HH, EE)]:
        for ft in (EE, HH):
            p = Problem((34, 19, 7), (True, False, True), boundft=boundft, ft=ft, with_mu=(ft == HH), real_mass=True)
            A_ref, _ = p.oracle_csc()
            A = p.operator(device=0, kernel=KERNELS["tiled"])
            x = p.random_x()
            err = rel(_apply_dev(A, x), A_ref.matvec(x))
            A.close()
            assert err < TOL, (boundft, ft, err)
    p = Problem((14, 11, 9), (True, True, False), real_mass=True, npml=2, omega=1.1)
    A_ref, _ = p.oracle_csc()
    b = A_ref.matvec(p.random_x(3))
    x_ref = spla.splu(A_ref.to_scipy()).solve(b)
    A = p.operator(device=0, kernel=KERNELS["tiled"])
    for method in ("bicgstab", "qmr"):
        xs, info = A.solve(torch.from_numpy(b).cuda(), method=method, rtol=1e-11, maxit=20000, check_every=10)
        assert rel(xs.cpu().numpy(), x_ref) < 1e-8, (method, info)
    A.close()


@pytest.mark.parametrize("boundft", list(itertools.product([EE, HH], repeat=3)))
def test_apply_all_boundft(boundft):
    """all 2^3 boundft choices, both formulations, on the tiled kernel (ARR = 0 / 1 / 2 instantiations)."""
    for isbloch in ((True, False, True), (False, True, False)):
        for ft in (EE, HH):
            p = Problem((9, 6, 7), isbloch, boundft, full_eps=(ft == EE), with_mu=True, ft=ft, full_mu=(ft == HH))
            A_ref, _ = p.oracle_csc()
            A = p.operator(device=0, kernel=KERNELS["tiled"])
            x = p.random_x()
            err = rel(_apply_dev(A, x), A_ref.matvec(x))
            errT = rel(_apply_dev(A, x, transpose=True), A_ref.to_scipy().T.tocsc() @ x)
            A.close()
            assert err < TOL and errT < TOL, (boundft, isbloch, ft, err, errT)


MIRRORED = (  # (ft, boundft, full_eps, with_mu): first curl backward on every axis
    (HH, (EE, EE, EE), False, True),     # A = Ce eps^-1 Cm - w^2 mu (model.jl:238-240), default boundft
    (HH, (EE, EE, EE), False, "full"),   # ... with a full 3x3 mu tensor as the mass parameter
    (EE, (HH, HH, HH), True, True),      # EE formulation on the dual arrangement, full tensor
    (EE, (HH, HH, HH), False, False),
)


@pytest.mark.parametrize("N", [(1, 1, 1), (2, 1, 3), (5, 3, 2), (31, 15, 4), (33, 17, 9), (70, 45, 6)])
def test_mirrored_arrangement_on_tiled_kernel(N):
    """N3 row: the HH formulation / boundft all-HH run on the mirrored variant of the tiled kernel (REV): every
    Bloch/symmetry combination, both layouts, forward and transposed."""
    for isbloch in itertools.product([True, False], repeat=3):
        for ft, boundft, full_eps, with_mu in MIRRORED:
            cmpfirst = not (isbloch[0] ^ isbloch[2])      # alternate the layout over the combinations
            p = Problem(N, isbloch, boundft, full_eps=full_eps, with_mu=bool(with_mu), ft=ft, cmpfirst=cmpfirst,
                        full_mu=(with_mu == "full"))
            A_ref, _ = p.oracle_csc()
            A = p.operator(device=0, kernel=KERNELS["tiled"])
            x = p.random_x()
            err = rel(_apply_dev(A, x), A_ref.matvec(x))
            errT = rel(_apply_dev(A, x, transpose=True), A_ref.to_scipy().T.tocsc() @ x)
            A.close()
            assert err < TOL and errT < TOL, (N, isbloch, ft, boundft, full_eps, cmpfirst, err, errT)


def test_mirrored_arrangement_deep_grid_and_sparse_offdiag():
    """several z-chunks per tile column (downward march), off-diagonal eps on part of the grid only (occupancy-mask
    path of the fused kernel), against the matrix-free oracle; tiled == general kernel."""
    torch = _torch()
    # z1 = 47: > 25 % of the (tile, plane) blocks flagged -> fused full-tensor kernel with the mask;
    # z1 = 33: sparse -> diagonal kernel + mirrored marching correction kernel
    for isbloch, z1, lo, hi in (((False, False, False), 47, 0.25, 0.6), ((True, False, True), 33, 0.0, 0.25)):
        p = Problem((70, 45, 90), isbloch, (HH, HH, HH), full_eps=True, with_mu=True)
        for v, u in itertools.permutations(range(3), 2):
            p.eps[:, :, :20, v, u] = 0
            p.eps[:, :, z1:, v, u] = 0
            p.eps[:25, :, :, v, u] = 0
        mf = p.oracle_matfree()
        x = p.random_x()
        At = p.operator(device=0, kernel=KERNELS["tiled"])
        An = p.operator(device=0, kernel=KERNELS["naive"])
        assert lo < At.offdiag_fraction < hi
        yt, yn = _apply_dev(At, x), _apply_dev(An, x)
        assert rel(yt, yn) < 1e-13
        assert rel(yt, mf(x)) < TOL
        At.close()
        An.close()
    p = Problem((40, 33, 70), (False, True, False), ft=HH, with_mu=True)
    A = p.operator(device=0, kernel=KERNELS["tiled"])
    x = p.random_x()
    assert rel(_apply_dev(A, x), p.oracle_matfree()(x)) < TOL
    A.close()


@pytest.mark.parametrize("boundft", [(EE, HH, EE), (HH, EE, EE), (HH, HH, EE), (EE, EE, HH), (EE, HH, HH), (HH, EE, HH)])
def test_mixed_boundft_on_tiled_kernel(boundft):
    """first curl forward on some axes only (ARR = 2: per-axis directions at run time): several z-chunks and tiles,
    sparse off-diagonal eps (diagonal kernel + marching correction) and dense (fused kernel), both layouts, HH
    formulation; against the matrix-free oracle and the general kernel."""
    for isbloch, z1, cmpfirst in (((True, False, True), 33, True), ((False, True, False), 60, False)):
        p = Problem((70, 45, 90), isbloch, boundft, full_eps=True, with_mu=True, cmpfirst=cmpfirst)
        for v, u in itertools.permutations(range(3), 2):
            p.eps[:, :, :20, v, u] = 0
            p.eps[:, :, z1:, v, u] = 0
            p.eps[:25, :, :, v, u] = 0
        x = p.random_x()
        At = p.operator(device=0, kernel=KERNELS["tiled"])
        An = p.operator(device=0, kernel=KERNELS["naive"])
        yt, yn = _apply_dev(At, x), _apply_dev(An, x)
        assert rel(yt, yn) < 1e-13, (boundft, isbloch)
        assert rel(yt, p.oracle_matfree()(x)) < TOL, (boundft, isbloch)
        At.close()
        An.close()
    p = Problem((40, 33, 50), (True, True, False), boundft, ft=HH, full_mu=True)
    A = p.operator(device=0, kernel=KERNELS["tiled"])
    x = p.random_x()
    assert rel(_apply_dev(A, x), p.oracle_matfree()(x)) < TOL
    A.close()


@pytest.mark.parametrize("method", ["bicgstab", "qmr"])
def test_solve_hh_formulation(method):
    """HH formulation end to end (fused-dot epilogue of the mirrored kernel inside BiCGSTAB): PML box, magnetic
    dipole; field vs sparse direct solve of the oracle matrix."""
    import scipy.sparse.linalg as spla
    fb = _fb()
    grid, (sdl_e, sdl_m, sei, smi), omega, eps, mu = _pml_box()
    ph = np.ones(3, complex)
    Ce, Cm = op.create_curls(sei, smi, (EE,) * 3, grid.isbloch, ph)
    Pe, Pm = op.create_paramops(eps, mu, sdl_e, sdl_m, sei, smi, (EE,) * 3, grid.isbloch, ph)
    A_ref = op.create_A(HH, omega, Pe, Pm, Ce, Cm)
    b = np.zeros(A_ref.shape[0], complex)
    n = grid.N[0]
    b[2 + 3 * ((n // 2) + n * ((n // 2) + n * (n // 2)))] = 1.0    # z-directed magnetic dipole at the centre
    h_ref = spla.splu(A_ref.to_scipy().tocsc()).solve(b)
    A = fb.FdfdOperator(grid.N, grid.isbloch, sdl_e, sdl_m, omega, eps, None, ph, ft="H", device=0, kernel=KERNELS["tiled"])
    x, info = A.solve(b, method=method, rtol=1e-10, maxit=20000, check_every=25)
    assert info["converged"], info
    assert rel(A_ref.matvec(x), b) < 1e-8
    assert rel(x, h_ref) < 1e-8 * 50
    A.close()


@pytest.mark.parametrize("kernel", ["tiled", "naive"])
def test_component_major_layout_and_weighted_out(kernel):
    for cmpfirst, wo in ((False, False), (True, True), (False, True)):
        p = Problem((37, 20, 11), (True, False, True), full_eps=True, with_mu=True, cmpfirst=cmpfirst, weighted_out=wo)
        A_ref, _ = p.oracle_csc()
        A = p.operator(device=0, kernel=KERNELS[kernel])
        x = p.random_x()
        err = rel(_apply_dev(A, x), A_ref.matvec(x))
        A.close()
        assert err < TOL, (cmpfirst, wo, err)


@pytest.mark.parametrize("kernel", ["tiled", "naive"])
def test_transpose_apply(kernel):
    for isbloch, full_eps, with_mu in (((True, True, True), True, True), ((False, True, False), True, False),
                                       ((False, False, False), False, True)):
        p = Problem((12, 35, 9), isbloch, full_eps=full_eps, with_mu=with_mu)
        A_ref, _ = p.oracle_csc()
        At = A_ref.to_scipy().T.tocsc()
        A = p.operator(device=0, kernel=KERNELS[kernel])
        x = p.random_x()
        err = rel(_apply_dev(A, x, transpose=True), At @ x)
        A.close()
        assert err < TOL, (isbloch, full_eps, err)


def test_omega_zero_skips_mass_term():
    p = Problem((10, 9, 8), (True, False, True), omega=0.0)
    A_ref, _ = p.oracle_csc()
    for k in (1, 2):
        A = p.operator(device=0, kernel=k)
        x = p.random_x()
        assert rel(_apply_dev(A, x), A_ref.matvec(x)) < TOL
        A.close()


def test_host_and_device_paths_agree_and_errors():
    fb = _fb()
    p = Problem((20, 18, 10), (False, True, False), full_eps=True)
    A = p.operator(device=0)
    x = p.random_x()
    yh = A @ x                       # FDFD_HOST
    yd = _apply_dev(A, x)            # FDFD_DEVICE
    assert np.array_equal(yh, yd)
    with pytest.raises(ValueError):
        A @ x[:-1]
    torch = _torch()
    xd = torch.from_numpy(x).cuda()
    with pytest.raises(fb._lib.FdfdError):      # aliasing is rejected
        A.mul(xd, xd)
    # changing omega re-scales the mass term
    A.set_omega(0.7)
    p2 = Problem((20, 18, 10), (False, True, False), full_eps=True, omega=0.7)
    assert rel(A @ x, p2.oracle_csc()[0].matvec(x)) < TOL
    A.close()


def test_export_pattern_on_gpu_handle_is_bit_exact():
    p = Problem((7, 6, 5), (False, True, False), full_eps=True, with_mu=True)
    A_ref, _ = p.oracle_csc()
    A = p.operator(device=0)
    cp, rv, nz = A.export_pattern()
    assert np.array_equal(cp, A_ref.julia_pattern()[0]) and np.array_equal(rv, A_ref.julia_pattern()[1])
    # the exported matrix and the matrix-free kernel are the same operator
    import scipy.sparse as sp
    S = sp.csc_matrix((nz, rv - 1, cp - 1), shape=A_ref.shape)
    x = p.random_x()
    assert rel(_apply_dev(A, x), S @ x) < TOL
    A.close()


def test_config_shapes_against_oracle():
    """BASELINE configs: C1 (40^3, CSC oracle) and reduced C2 / C3 / C4 / C5 (matrix-free oracle), real inputs."""
    import workloads
    from oracle.matfree import MatFreeOperator
    rng = np.random.default_rng(SEED)
    for w in (workloads.c1_vacuum_box(), workloads.c2_waveguide((72, 60, 40)), workloads.c3_phc_slab((64, 64, 40)),
              workloads.c4_scatterer((48, 48, 48), radius_cells=12), workloads.c5_metalens((64, 64, 96))):
        A = workloads.make_operator(w, device=0)
        eps = w["eps"]
        if w.get("julia_layout"):       # [u,v,k,j,i] (the memory of Julia's (Nx,Ny,Nz,3,3) array) -> [i,j,k,v,u]
            eps = np.ascontiguousarray(np.transpose(eps, (4, 3, 2, 1, 0)))
        mf = MatFreeOperator(EE, w["omega"], eps, None, w["sdl_e"], w["sdl_m"], (EE,) * 3, w["isbloch"], w["e_mikL"])
        x = crandn(rng, A.n)
        err = rel(_apply_dev(A, x), mf(x))
        A.close()
        assert err < TOL, (w["name"], err)


def test_golden_fixture():
    """committed oracle output (tests/golden/make_golden.py) - guards against oracle AND kernel drift."""
    path = os.path.join(os.path.dirname(__file__), "golden", "apply_golden.npz")
    g = np.load(path)
    for tag in ("bloch_full", "sym_diag"):
        kw = dict(N=tuple(g[f"{tag}_N"]), isbloch=tuple(bool(b) for b in g[f"{tag}_isbloch"]),
                  full_eps=bool(g[f"{tag}_full"]), with_mu=bool(g[f"{tag}_mu"]))
        p = Problem(**kw)
        x = p.random_x()
        assert np.array_equal(x, g[f"{tag}_x"])
        for k in (1, 2):
            A = p.operator(device=0, kernel=k)
            assert rel(_apply_dev(A, x), g[f"{tag}_y"]) < TOL
            cp, rv, _ = A.export_pattern(values=False)
            assert np.array_equal(cp, g[f"{tag}_colptr"]) and np.array_equal(rv, g[f"{tag}_rowval"])
            A.close()


def test_large_grid_properties():
    """full-size-style properties where the CSC oracle does not fit: tiled == general kernel, linearity,
    and agreement with the matrix-free oracle on a z-sub-slab."""
    import workloads
    from oracle.matfree import MatFreeOperator
    torch = _torch()
    w = workloads.c2_waveguide((200, 200, 48))
    At = workloads.make_operator(w, device=0, kernel=2)
    An = workloads.make_operator(w, device=0, kernel=1)
    g = torch.Generator(device="cuda").manual_seed(3)
    x1 = torch.randn(At.n, 2, device="cuda", dtype=torch.float64, generator=g).view(torch.complex128).reshape(-1)
    x2 = torch.randn(At.n, 2, device="cuda", dtype=torch.float64, generator=g).view(torch.complex128).reshape(-1)
    y1, y2 = At @ x1, At @ x2
    yn = An @ x1
    assert float((y1 - yn).norm() / yn.norm()) < 1e-13
    a = 0.3 - 1.7j
    y12 = At @ (x1 + a * x2)
    assert float((y12 - (y1 + a * y2)).norm() / y12.norm()) < 1e-13
    mf = MatFreeOperator(EE, w["omega"], w["eps"], None, w["sdl_e"], w["sdl_m"], (EE,) * 3, w["isbloch"], w["e_mikL"])
    assert rel(y1.cpu().numpy(), mf(x1.cpu().numpy())) < TOL
    At.close()
    An.close()


# ---------------------------------------------------------------------------------------------------
# solve / RHS / post-processing
# ---------------------------------------------------------------------------------------------------
def _pml_box(n=20, npml=5):
    from oracle.grid import Grid, create_stretched_dls
    d, lam = 1.0, 10.0
    omega = 2 * np.pi / lam
    lp = (np.arange(n + 1) - n / 2) * d
    grid = Grid((lp, lp, lp), (False, False, False))
    sdl = create_stretched_dls(omega, grid, ((npml,) * 3, (npml,) * 3))
    N = grid.N
    eps = np.zeros(N + (3, 3), complex)
    mu = np.zeros(N + (3, 3), complex)
    for v in range(3):
        eps[..., v, v] = 1 + 0.5 * (np.abs(np.arange(n) - n / 2)[:, None, None] < 3)
        mu[..., v, v] = 1
    return grid, sdl, omega, eps, mu


@pytest.mark.parametrize("method", ["bicgstab", "qmr"])
def test_solve_matches_direct_solve(method):
    import scipy.sparse.linalg as spla
    from oracle.source import PointSrc, add_src, create_field_array
    fb = _fb()
    grid, (sdl_e, sdl_m, sei, smi), omega, eps, mu = _pml_box()
    ph = np.ones(3, complex)
    Ce, Cm = op.create_curls(sei, smi, (EE,) * 3, grid.isbloch, ph)
    Pe, Pm = op.create_paramops(eps, mu, sdl_e, sdl_m, sei, smi, (EE,) * 3, grid.isbloch, ph)
    A_ref = op.create_A(EE, omega, Pe, Pm, Ce, Cm)
    je = create_field_array(grid.N)
    add_src(je, EE, (EE,) * 3, grid, PointSrc([0.3, 0.2, 0.1], [0, 0, 1]))
    b = op.create_b(EE, omega, Pe, Pm, Ce, Cm, op.field_arr2vec(je), np.zeros(A_ref.shape[0]))
    e_ref = spla.splu(A_ref.to_scipy().tocsc()).solve(b)
    A = fb.FdfdOperator(grid.N, grid.isbloch, sdl_e, sdl_m, omega, eps, None, ph, device=0)
    # RHS through the GPU create_b equals the oracle's
    assert rel(A.create_b(op.field_arr2vec(je)), b) < TOL
    torch = _torch()
    x, info = A.solve(torch.from_numpy(b).cuda(), method=method, rtol=1e-10, maxit=20000, check_every=25, history=True)
    assert info["converged"], info
    e = x.cpu().numpy()
    assert rel(A_ref.matvec(e), b) < 1e-8                       # true residual
    assert rel(e, e_ref) < 1e-8 * 50                            # field vs direct solve (cond. number slack)
    h = info["history"]
    assert h[0] == pytest.approx(1.0) and np.isfinite(h).all() and h[-1] <= 1e-10
    # post-processing: h_from_e
    h_ref = op.h_from_e(e_ref, omega, Pm, Ce, np.zeros_like(e_ref))
    assert rel(A.h_from_e(e_ref), h_ref) < TOL
    # host-buffer solve gives the same answer
    xh, info_h = A.solve(b, method=method, rtol=1e-10, maxit=20000, check_every=25)
    assert info_h["converged"] and rel(xh, e) < 1e-6
    A.close()


def test_solve_edge_cases():
    fb = _fb()
    p = Problem((8, 7, 6), (True, True, True))
    A = p.operator(device=0)
    x, info = A.solve(np.zeros(A.n, complex))                    # b == 0
    assert info["converged"] and info["iters"] == 0 and not x.any()
    b = p.oracle_csc()[0].matvec(p.random_x(5))
    x, info = A.solve(b, maxit=3, rtol=1e-14)                    # maxit hit: ENOCONV, x still valid
    assert not info["converged"] and info["iters"] == 3 and np.isfinite(x).all()
    A.close()
    # exact convergence between two residual checks (3 unknowns, check every 10 iterations): the iteration idles on
    # the solution instead of dividing 0 by 0
    p1 = Problem((1, 1, 1), (True, True, True), kb_scale=3.0)
    A = p1.operator(device=0)
    A_ref = p1.oracle_csc()[0]
    b = A_ref.matvec(p1.random_x(2))
    x, info = A.solve(b, rtol=1e-12, maxit=50, check_every=10)
    assert info["converged"] and np.isfinite(x).all() and rel(A_ref.matvec(x), b) < 1e-12, info
    A.close()


def test_create_b_with_magnetic_current_and_h_from_e_with_mu():
    p = Problem((9, 8, 7), (True, False, True), with_mu=True)
    A_ref, (Pe, Pm, Ce, Cm) = p.oracle_csc()
    A = p.operator(device=0)
    je, jm, e = p.random_x(11), p.random_x(12), p.random_x(13)
    assert rel(A.create_b(je, jm), op.create_b(EE, p.omega, Pe, Pm, Ce, Cm, je, jm)) < TOL
    assert rel(A.h_from_e(e, jm), op.h_from_e(e, p.omega, Pm, Ce, jm)) < TOL
    A.close()


def test_e_from_h_and_corner_interpolation():
    """N2 rows: e_from_h (model.jl:281-284) and the create_Mcs operators (model.jl:287-306) vs the oracle."""
    for isbloch, boundft in (((True, False, True), (EE, EE, EE)), ((False, True, False), (EE, EE, EE)),
                             ((True, True, False), (HH, EE, HH))):
        p = Problem((9, 8, 7), isbloch, boundft, with_mu=True)
        A_ref, (Pe, Pm, Ce, Cm) = p.oracle_csc()
        A = p.operator(device=0)
        h, je, e = p.random_x(21), p.random_x(22), p.random_x(23)
        assert rel(A.e_from_h(h, je), op.e_from_h(h, p.omega, Pe, Cm, je)) < TOL
        assert rel(A.e_from_h(h), op.e_from_h(h, p.omega, Pe, Cm, np.zeros_like(h))) < TOL
        Mce, Mcm = op.create_Mcs(p.sdl_e, p.sdl_m, p.sei, p.smi, boundft, isbloch, p.ph)
        assert rel(A.interp_corners(e, "E"), Mce.matvec(e)) < TOL
        assert rel(A.interp_corners(h, "H"), Mcm.matvec(h)) < TOL
        # round trip e -> h -> e with consistent sources: e_from_h(h_from_e(e)) solves the same Maxwell pair
        A.close()
    fb = _fb()
    p = Problem((6, 5, 4), full_eps=True)
    A = p.operator(device=0)
    with pytest.raises(fb._lib.FdfdError):      # Peps must be diagonal for e_from_h (reference `Peps \`)
        A.e_from_h(p.random_x())
    A.close()


def test_tfsf_rhs_reproduces_the_incident_wave():
    """TF/SF stand-in for C4 (workloads.tfsf_rhs): b = A (M E_inc) - M (A E_inc) lives on the surface of the
    total-field box only, and in vacuum the solve returns exactly M E_inc (the discrete plane wave inside the box,
    nothing outside)."""
    import workloads
    w = workloads.c4_scatterer((40, 40, 40), radius_cells=-5)       # negative radius: vacuum everywhere
    A = workloads.make_operator(w, device=0)
    b, e_ref = workloads.tfsf_rhs(A, w, box_cells=16)
    B = np.abs(_fb().field_vec2arr(b, w["N"])).max(axis=-1)
    assert B[20, 20, 20] == 0 and B[2, 2, 2] == 0                   # deep inside / far outside: exactly zero
    assert B[20, 20, 28] > 0 and B[20, 20, 12] > 0                  # the two z faces of the box
    assert np.count_nonzero(B) < 0.05 * B.size
    e, info = A.solve(b, rtol=1e-10, maxit=20000, check_every=50)
    assert info["converged"]
    assert np.abs(e - e_ref).max() < 1e-5                            # |E_inc| = 1
    A.close()


def test_rhs_and_postprocessing_on_hh_handles():
    """create_b (HH branch, model.jl:267-270), h_from_e and e_from_h evaluated on an FT_HH handle vs the oracle."""
    for isbloch, boundft in (((True, False, True), (EE, EE, EE)), ((False, True, False), (HH, EE, HH))):
        p = Problem((9, 8, 7), isbloch, boundft, with_mu=True, ft=HH)
        A_ref, (Pe, Pm, Ce, Cm) = p.oracle_csc()
        A = p.operator(device=0)
        je, jm, e, h = p.random_x(31), p.random_x(32), p.random_x(33), p.random_x(34)
        assert rel(A.create_b(je, jm), op.create_b(HH, p.omega, Pe, Pm, Ce, Cm, je, jm)) < TOL
        assert rel(A.create_b(je), op.create_b(HH, p.omega, Pe, Pm, Ce, Cm, je, np.zeros_like(je))) < TOL
        assert rel(A.e_from_h(h, je), op.e_from_h(h, p.omega, Pe, Cm, je)) < TOL
        assert rel(A.h_from_e(e, jm), op.h_from_e(e, p.omega, Pm, Ce, jm)) < TOL
        assert rel(A.h_from_e(e), op.h_from_e(e, p.omega, Pm, Ce, np.zeros_like(e))) < TOL
        A.close()
    # identity mu (scalar mass parameter) on the HH handle
    p = Problem((8, 7, 6), (True, True, False), ft=HH)
    A_ref, (Pe, Pm, Ce, Cm) = p.oracle_csc()
    A = p.operator(device=0)
    e, jm = p.random_x(41), p.random_x(42)
    assert rel(A.h_from_e(e, jm), op.h_from_e(e, p.omega, Pm, Ce, jm)) < TOL
    A.close()


@pytest.mark.parametrize("ft", [EE, HH])
def test_reference_call_sequence(ft):
    """The reference's own sequence (model.jl:141-284): Ps = create_paramops(mdl); Cs = create_curls(mdl);
    js = create_srcs(mdl); A, b = create_linsys(ft, w, Ps, Cs, js); solve; h_from_e / e_from_h(.., w, Ps, Cs, js) -
    every piece against the oracle built from the same model inputs."""
    from oracle.grid import Grid as OGrid, create_stretched_dls as o_sdls
    fb = _fb()
    n = 14
    lp = (np.arange(n + 1) - n / 2) * 1.0
    mdl = fb.ModelFull(fb.Grid((lp, lp, lp), (False, False, False)))
    w = 2 * np.pi / 8.0
    fb.set_wpml(mdl, w)
    fb.set_Npml(mdl, ((3,) * 3, (3,) * 3))
    rng = np.random.default_rng(SEED)
    for v in range(3):
        mdl.eps_arr[..., v, v] = 1.0 + 0.5 * rng.random(mdl.grid.N)
        mdl.mu_arr[..., v, v] = 1.0 + 0.2 * rng.random(mdl.grid.N)
    fb.add_srce(mdl, fb.PointSrc([0.3, 0.2, 0.1], [0, 0, 1]))
    fb.add_srcm(mdl, fb.PointSrc([-1.2, 0.4, 0.6], [1, 0, 0]))
    Ps, Cs, js = fb.create_paramops(mdl), fb.create_curls(mdl), fb.create_srcs(mdl)
    A, b = fb.create_linsys(ft, w, Ps, Cs, js, device=0)
    assert fb.create_A(ft, w, Ps, Cs, device=0) is A                  # one operator per (formulation, w)
    # oracle from the same inputs
    og = OGrid((lp, lp, lp), (False, False, False))
    sdl_e, sdl_m, sei, smi = o_sdls(w, og, ((3,) * 3, (3,) * 3))
    ph = np.ones(3, complex)
    Ce, Cm = op.create_curls(sei, smi, (EE,) * 3, og.isbloch, ph)
    Pe, Pm = op.create_paramops(mdl.eps_arr, mdl.mu_arr, sdl_e, sdl_m, sei, smi, (EE,) * 3, og.isbloch, ph)
    A_ref = op.create_A(ft, w, Pe, Pm, Ce, Cm)
    assert rel(b, op.create_b(ft, w, Pe, Pm, Ce, Cm, js[0], js[1])) < TOL
    assert rel(fb.create_b(ft, w, Ps, Cs, js, device=0), b) == 0.0
    x, info = fb.solve(A, b, rtol=1e-10, maxit=20000)
    assert info["converged"] and rel(A_ref.matvec(x), b) < 1e-8
    if ft == EE:
        assert rel(fb.h_from_e(x, w, Ps, Cs, js, device=0), op.h_from_e(x, w, Pm, Ce, js[1])) < TOL
    else:
        assert rel(fb.e_from_h(x, w, Ps, Cs, js, device=0), op.e_from_h(x, w, Pe, Cm, js[0])) < TOL
    A.close()


def test_model_api_end_to_end():
    """reference-shaped host API: ModelFull -> add_srce -> create_linsys -> solve -> h_from_e."""
    fb = _fb()
    n = 16
    lp = (np.arange(n + 1) - n / 2) * 1.0
    mdl = fb.ModelFull(fb.Grid((lp, lp, lp), (False, False, False)))
    w = 2 * np.pi / 8.0
    fb.set_wpml(mdl, w)
    fb.set_Npml(mdl, ((4,) * 3, (4,) * 3))
    for v in range(3):
        mdl.eps_arr[..., v, v] = 1.0
    fb.add_srce(mdl, fb.PointSrc([0.0, 0.0, 0.0], [0, 0, 1]))
    A, b = fb.create_linsys(fb.EE, w, mdl)
    e, info = fb.solve(A, b, rtol=1e-9, maxit=5000)
    assert info["converged"]
    assert rel(A @ e, b) < 1e-8
    h = fb.h_from_e(e, w, A)
    assert np.isfinite(h).all() and np.abs(h).max() > 0
    A.close()


def test_symmetric_offdiagonal_tensor_is_stored_once():
    """P_vu == P_uv pointwise (what subpixel smoothing of reciprocal media gives): three off-diagonal arrays, the
    other three slots alias them - same results from every kernel path, forward and transposed, both formulations."""
    for N, isbloch, cmpfirst, ft in (((37, 20, 11), (True, False, True), True, EE), ((33, 9, 40), (False, True, False), False, EE),
                                     ((20, 13, 9), (True, True, True), True, HH)):
        for sym in (True, False):
            p = Problem(N, isbloch, full_eps=(ft == EE), full_mu=(ft == HH), with_mu=True, cmpfirst=cmpfirst, ft=ft)
            m = p.eps if ft == EE else p.mu
            if sym:
                for v, u in itertools.combinations(range(3), 2):
                    m[..., u, v] = m[..., v, u]
            A_ref, _ = p.oracle_csc()
            x = p.random_x()
            for k in (2, 1):
                A = p.operator(device=0, kernel=k)
                assert A.offdiag_symmetric == sym
                assert rel(_apply_dev(A, x), A_ref.matvec(x)) < TOL
                assert rel(_apply_dev(A, x, transpose=True), A_ref.to_scipy().T.tocsc() @ x) < TOL
                A.close()
            for v, u in itertools.permutations(range(3), 2):      # sparse pattern: diagonal kernel + correction pass
                m[:, :, 2:, v, u] = 0
            A = p.operator(device=0, kernel=2)
            assert rel(_apply_dev(A, x), p.oracle_csc()[0].matvec(x)) < TOL
            A.close()


# ---------------------------------------------------------------------------------------------------
# N4: material pipeline (kept last in this file: the newest kernel)
# ---------------------------------------------------------------------------------------------------
# ---- written after the round-1 GPU budget was spent (checked under the CPU logic-check build only): kept last so that
# a surprise on hardware cannot hide the results of the tests above behind `pytest -x` ------------------------------
@pytest.mark.parametrize("classic", [False, True])
@pytest.mark.parametrize("ft,full", [(EE, True), (EE, False), (HH, False)])
def test_bicgstab_trajectory_matches_textbook_iteration(ft, full, classic):
    """Inner products that are wrong (sigma accumulated by the p update from u = A^T conj(rhat), (t,s) / (t,t) from the
    apply epilogue and the correction-pass deltas) keep r = b - A x consistent and may still converge: only the
    iterates themselves show them.  Five iterations against the textbook BiCGSTAB on the oracle operator, both with
    the default schedule and with the separate (rhat, v) pass (FDFD_BICGSTAB_CLASSIC)."""
    p = Problem((33, 20, 11), (True, False, True), ft=ft, omega=1.2 - 0.3j, full_eps=full and ft == EE, with_mu=ft == HH)
    mf = p.oracle_matfree()
    b = p.random_x(3)
    x_ref = np.zeros_like(b)                                    # textbook iteration (van der Vorst), numpy
    r = b.copy(); rh = r.copy(); pv = r.copy(); rho = np.vdot(rh, r)
    for _ in range(5):
        v = mf(pv)
        alpha = rho / np.vdot(rh, v)
        sv = r - alpha * v
        t = mf(sv)
        omega = np.vdot(t, sv) / np.vdot(t, t)
        x_ref = x_ref + alpha * pv + omega * sv
        r = sv - omega * t
        rho_new = np.vdot(rh, r)
        pv = r + (rho_new / rho) * (alpha / omega) * (pv - omega * v)
        rho = rho_new
    A = p.operator(device=0, kernel=KERNELS["tiled"])
    old = os.environ.pop("FDFD_BICGSTAB_CLASSIC", None)
    try:
        if classic:
            os.environ["FDFD_BICGSTAB_CLASSIC"] = "1"
        x, info = A.solve(b, method="bicgstab", rtol=1e-300, maxit=5, check_every=1)
    finally:
        os.environ.pop("FDFD_BICGSTAB_CLASSIC", None)
        if old is not None:
            os.environ["FDFD_BICGSTAB_CLASSIC"] = old
    assert info["iters"] == 5
    assert rel(x, x_ref) < 1e-9
    assert abs(rel(mf(x), b) - info["relres"]) < 1e-9
    A.close()


def _reduced_cases():
    from problems import REDUCED_CASES
    return REDUCED_CASES


@pytest.mark.parametrize("case", _reduced_cases(), ids=lambda c: f"{c[0]}-{'x'.join(map(str, c[1]))}-ft{c[4]}")
def test_reduced_models_match_k_dimensional_oracle(case):
    """ModelTE / ModelTM / ModelTEM (te.jl, tm.jl, tem.jl) run on the 3-D kernels (one periodic cell along the missing
    axes); the reference call sequence against the operators assembled on the K-dimensional grid (oracle/reduced.py):
    A x, A^T x, b, h_from_e / e_from_h within 1e-12, solved field within 1e-8 * cond. slack"""
    from problems import reduced_model_check
    errs = reduced_model_check(_fb(), *case)
    assert max(v for k, v in errs.items() if k != "solve") < TOL, errs
    assert errs["solve"] < 1e-8 * 50, errs


def test_reduced_operator_on_device_tensors():
    """ReducedOperator with torch CUDA tensors: embedding / extraction on the device, same numbers as the host path"""
    torch = _torch()
    fb = _fb()
    rng = np.random.default_rng(1)
    for cmpfirst in (True, False):
        mdl = fb.ModelTE(fb.Grid([np.arange(8.0), np.arange(6.0)], (True, False)))
        mdl.order_cmpfirst = cmpfirst
        mdl.eps_arr[..., 0, 0], mdl.eps_arr[..., 1, 1] = 2.0, 3.0
        mdl.eps_arr[..., 0, 1] = mdl.eps_arr[..., 1, 0] = 0.1
        A = fb.create_A(EE, 1.0, mdl, device=0)
        x = crandn(rng, A.n)
        y = A @ x
        yd = A @ torch.from_numpy(x).cuda()
        assert yd.is_cuda and rel(yd.cpu().numpy(), y) < 1e-14
        xs, info = A.solve(torch.from_numpy(y).cuda(), rtol=1e-12)
        assert info["converged"] and rel(xs.cpu().numpy(), x) < 1e-9
        A.close()


# ---- material pipeline (a kernel that has not run on hardware yet): last of all ---------------------------------------
def test_calc_matparams_matches_oracle():
    """fdfd_calc_matparams (object assignment + Kottke smoothing, one kernel) vs oracle/matparams.py: boxes, balls,
    cylinders, Bloch / symmetry ghost corners, non-uniform grids, full-tensor materials, mu locations, z-slabs.
    Tolerance 1e-10 relative (fp64; measured 1e-13..1e-14 under the CPU logic-check build): the plane-cut volume
    formula cancels for nearly axis-parallel normals, which amplifies FMA-contraction differences."""
    from oracle import matparams as omp
    from oracle.grid import Grid as OGrid
    from problems import matparams_scene, MATPARAMS_CASES
    fb = _fb()
    golden = np.load(os.path.join(os.path.dirname(__file__), "golden", "matparams_golden.npz"))
    for case_no, (N, isbloch, boundft, ft, uniform, nshape, aniso) in enumerate(MATPARAMS_CASES):
        lp, o_sh, f_sh, pinds, params = matparams_scene(N, isbloch, uniform, nshape, aniso)
        ref = omp.calc_matparams(OGrid(lp, isbloch), boundft, ft, o_sh, pinds, params)
        g = fb.Grid(lp, isbloch)
        got = fb.calc_matparams_array(g, boundft, ft, f_sh, pinds, params, device=0)
        assert rel(got, ref) < 1e-10, (N, isbloch, boundft, ft, rel(got, ref))
        if f"case{case_no}" in golden:                       # committed fixture (tests/golden/make_golden_matparams.py)
            assert rel(got, golden[f"case{case_no}"]) < 1e-10
        slab = fb.calc_matparams_array(g, boundft, ft, f_sh, pinds, params, k0=2, k1=N[2] - 1, device=0)
        assert np.array_equal(slab, got[:, :, 2:N[2] - 1])
    with pytest.raises(fb._lib.FdfdError):
        fb.calc_matparams_array(g, boundft, ft, f_sh[1:2], [0], params[:1], device=0)


def test_model_with_objects_end_to_end():
    """reference sequence with objects: add_obj! -> create_paramops (calc_matparams!, model.jl:143) -> create_A;
    the operator built from the GPU-smoothed arrays equals the oracle operator built from the oracle-smoothed arrays."""
    from oracle import matparams as omp
    from oracle.grid import Grid as OGrid, create_stretched_dls as o_sdls
    fb = _fb()
    n = 12
    lp = (np.arange(n + 1) - n / 2) * 1.0
    mdl = fb.ModelFull(fb.Grid((lp, lp, lp), (False, False, False)))
    w = 2 * np.pi / 8.0
    fb.set_wpml(mdl, w)
    fb.set_Npml(mdl, ((2,) * 3, (2,) * 3))
    fb.add_obj(mdl, "vacuum", fb.Box([0, 0, 0], [10, 10, 10]), eps=1.0)
    fb.add_obj(mdl, "glass", fb.Ball([0.3, -0.2, 0.1], 3.4), fb.Cylinder([-2, 2, 0], 1.5, 4.0, axis=0), eps=2.25)
    Ps, Cs = fb.create_paramops(mdl, device=0), fb.create_curls(mdl)
    o_sh = [omp.Box([0, 0, 0], [10, 10, 10]), omp.Ball([0.3, -0.2, 0.1], 3.4), omp.Cylinder([-2, 2, 0], 1.5, 4.0, 0)]
    og = OGrid((lp, lp, lp), (False, False, False))
    eps_ref = omp.calc_matparams(og, (EE,) * 3, EE, o_sh, [0, 1, 1], [np.eye(3), 2.25 * np.eye(3)])
    assert rel(mdl.eps_arr, eps_ref) < 1e-10
    assert np.abs(mdl.eps_arr[..., 0, 1]).max() > 1e-3          # the smoothing produced off-diagonal entries
    A = fb.create_A(fb.EE, w, Ps, Cs, device=0)
    sdl_e, sdl_m, sei, smi = o_sdls(w, og, ((2,) * 3, (2,) * 3))
    ph = np.ones(3, complex)
    mu = np.zeros(og.N + (3, 3), complex)
    for v in range(3):
        mu[..., v, v] = 1
    Ce, Cm = op.create_curls(sei, smi, (EE,) * 3, og.isbloch, ph)
    Pe, Pm = op.create_paramops(eps_ref, mu, sdl_e, sdl_m, sei, smi, (EE,) * 3, og.isbloch, ph)
    A_ref = op.create_A(EE, w, Pe, Pm, Ce, Cm)
    x = crandn(np.random.default_rng(SEED), A.n)
    assert rel(A @ x, A_ref.matvec(x)) < 1e-10
    A.close()


def test_eps_from_objects_on_the_device():
    """fdfd_set_eps_objects: the operator rasterises and smooths its own slab on the device (no host eps array); it is
    bit-for-bit the operator that is fed the array fdfd_calc_matparams returns, and symmetric materials are stored once"""
    from problems import matparams_scene
    fb = _fb()
    for N, isbloch, boundft, sym in (((12, 10, 9), (True, False, True), (EE, EE, EE), True),
                                     ((13, 9, 11), (False, True, False), (EE, HH, EE), False)):
        lp, _, f_sh, pinds, params = matparams_scene(N, isbloch, False, 6, not sym)
        p = Problem(N, isbloch, boundft)
        eps = fb.calc_matparams_array(fb.Grid(lp, isbloch), boundft, EE, f_sh, pinds, params, device=0)
        bf = ["E" if b == EE else "H" for b in boundft]
        A1 = fb.FdfdOperator(N, isbloch, p.sdl_e, p.sdl_m, p.omega, eps, None, p.ph, boundft=bf, device=0)
        A2 = fb.FdfdOperator(N, isbloch, p.sdl_e, p.sdl_m, p.omega, None, None, p.ph, boundft=bf, device=0)
        A2.set_eps_objects(lp, f_sh, pinds, params, boundft=bf)
        x = p.random_x()
        assert np.array_equal(_apply_dev(A1, x), _apply_dev(A2, x))
        assert A1.offdiag_symmetric == sym and A2.offdiag_symmetric == sym
        with pytest.raises(fb._lib.FdfdError):
            A2.export_pattern()
        A1.close()
        A2.close()
    # model API: objects -> operator without mdl.eps_arr ever being filled
    n = 12
    lpm = (np.arange(n + 1) - n / 2) * 1.0
    mdl = fb.ModelFull(fb.Grid((lpm, lpm, lpm), (False, False, False)))
    w = 2 * np.pi / 8.0
    fb.set_wpml(mdl, w)
    fb.set_Npml(mdl, ((2,) * 3, (2,) * 3))
    fb.add_obj(mdl, "vacuum", fb.Box([0, 0, 0], [10, 10, 10]), eps=1.0)
    fb.add_obj(mdl, "glass", fb.Ball([0.3, -0.2, 0.1], 3.4), eps=2.25)
    Ad = fb.create_A(fb.EE, w, fb.create_paramops(mdl, device=0, device_materials=True), fb.create_curls(mdl), device=0)
    assert not mdl.eps_arr.any()
    Ah = fb.create_A(fb.EE, w, fb.create_paramops(mdl, device=0), fb.create_curls(mdl), device=0)
    x = crandn(np.random.default_rng(SEED), Ad.n)
    assert np.array_equal(Ad @ x, Ah @ x)
    Ad.close()
    Ah.close()


def test_objects_on_reduced_models():
    """add_obj / calc_matparams on ModelTE, ModelTM, ModelTEM (te.jl:17-64, tm.jl:17-64, tem.jl:16-59): the material
    kernel on the extruded scene; analytic harmonic / arithmetic means across a planar interface, and the oracle's
    smoothing of the scene the K-dimensional one stands for (1e-10, as for the 3-D pipeline)"""
    from problems import reduced_objects_check
    assert reduced_objects_check(_fb()) == 22


@pytest.mark.parametrize("N", [(3, 3, 2), (33, 17, 9), (70, 45, 6), (31, 40, 5), (61, 29, 7)])
def test_apply_fused_full_tensor_rowpair(N):
    """Fused full-tensor shape of the row-pair kernel (symmetric tensor, real entries: diagonal and off-diagonal rows travel
    as doubles through the TMA ring, apply_rowpair.cu HAS_OFF) against the oracle's CSC product, every Bloch / symmetry
    combination (Bloch wrap of the forward neighbours at tile edges), forward and transposed.  The operator being
    matched: create_paramops' Mout * P_off * Min, /root/reference/src/model/model.jl:149-153."""
    for isbloch in itertools.product([True, False], repeat=3):
        p = Problem(N, isbloch, full_eps=True, real_mass=True, sym_real_off=True)
        A_ref, _ = p.oracle_csc()
        A = p.operator(device=0, kernel=KERNELS["tiled"])
        assert A.offdiag_bytes_per_dof == 8.0, "the fused row-pair shape was not selected"
        x = p.random_x()
        err = rel(_apply_dev(A, x), A_ref.matvec(x))
        errT = rel(_apply_dev(A, x, transpose=True), A_ref.to_scipy().T.tocsc() @ x)
        A.close()
        assert err < TOL and errT < TOL, (N, isbloch, err, errT)


def test_fused_full_tensor_partly_empty_blocks_arrangements_and_solve():
    """Fused row-pair shape on a deeper grid (several z-chunks and items per CTA) whose off-diagonal entries vanish on
    whole planes / half-spaces (tile occupancy mask, per-row-pair skip), on the default, mirrored and a mixed
    arrangement and on the HH formulation, against the general kernel (1e-13) and the matrix-free oracle; then BiCGSTAB
    (fused dots in the apply epilogue) and QMR (transposed operator) against a sparse direct solve."""
    import scipy.sparse.linalg as spla
    torch = _torch()
    cases = [((EE, EE, EE), (True, True, True), EE), ((HH, HH, HH), (False, True, False), EE),
             ((EE, HH, EE), (True, False, True), EE), ((EE, EE, EE), (False, False, True), HH)]
    for boundft, isbloch, ft in cases:
        p = Problem((64, 45, 40), isbloch, boundft, full_eps=(ft == EE), full_mu=(ft == HH), with_mu=(ft == HH), ft=ft,
                    real_mass=True, sym_real_off=True)
        mass = p.eps if ft == EE else p.mu
        for v, u in itertools.permutations(range(3), 2):
            mass[:, :, :9, v, u] = 0
            mass[:, :, 21:30, v, u] = 0
            mass[:, 11:30, 30:, v, u] = 0
            mass[:33, :, 12:18, v, u] = 0
        x = p.random_x()
        A = p.operator(device=0, kernel=KERNELS["tiled"])
        An = p.operator(device=0, kernel=KERNELS["naive"])
        assert A.offdiag_bytes_per_dof == 8.0 and 0.0 < A.offdiag_fraction < 1.0
        y = _apply_dev(A, x)
        e1, e2 = rel(y, _apply_dev(An, x)), rel(y, p.oracle_matfree()(x))
        A.close()
        An.close()
        assert e1 < 1e-13 and e2 < TOL, (boundft, isbloch, ft, e1, e2)
    p = Problem((14, 11, 9), (True, True, False), full_eps=True, real_mass=True, sym_real_off=True, npml=2, omega=1.1)
    A_ref, _ = p.oracle_csc()
    b = A_ref.matvec(p.random_x(3))
    x_ref = spla.splu(A_ref.to_scipy()).solve(b)
    A = p.operator(device=0, kernel=KERNELS["tiled"])
    for method in ("bicgstab", "qmr"):
        xs, info = A.solve(torch.from_numpy(b).cuda(), method=method, rtol=1e-11, maxit=20000, check_every=10)
        assert rel(xs.cpu().numpy(), x_ref) < 1e-8, (method, info)
    A.close()



