"""Property tests pinning the operator oracle.

The reference has no golden vectors for create_curls/create_paramops/create_A ("parity
unpinned"), so the restatement (oracle/operators.py, oracle/matfree.py) is pinned by
mathematical identities instead, all in fp64 to ~1e-13 relative.
"""
import itertools

import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle.grid import Grid, EE, HH, PRIM, DUAL, create_stretched_dls, create_e_mikL
from oracle import operators as op
from oracle.matfree import MatFreeOperator
from oracle.source import PlaneSrc, PointSrc, add_src, create_field_array

RNG = np.random.default_rng(20261017)


def crandn(*shape):
    return RNG.standard_normal(shape) + 1j * RNG.standard_normal(shape)


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def dense(A):
    return A.to_scipy().toarray()


# ---- P1 -------------------------------------------------------------------------------
@pytest.mark.parametrize("isbloch", [True, False])
def test_P1_adjoint_of_differences(isbloch):
    N = (5, 6, 7)
    for w in range(3):
        ph = np.exp(-0.3j * (w + 1)) if isbloch else 1.0
        Dp = dense(op.create_d(w, True, N, None, isbloch, ph))
        Dm = dense(op.create_d(w, False, N, None, isbloch, ph))
        if isbloch:
            assert np.abs(Dm + Dp.conj().T).max() < 1e-15
        else:
            # symmetry BC: D- = -(D+)^T away from the rows/cols the boundary rule zeroes
            idx = np.arange(np.prod(N)).reshape(N, order="F")
            sl = [slice(None)] * 3
            sl[w] = slice(1, None)
            inner = idx[tuple(sl)].ravel()
            assert np.abs((Dm + Dp.T)[np.ix_(inner, inner)]).max() < 1e-15
            first = np.take(idx, 0, axis=w).ravel()
            assert not Dm[first, :].any()          # row 1 of backward difference is zero
            assert not Dp[:, first].any()          # forward difference never reads f[1]


# ---- P2 -------------------------------------------------------------------------------
@pytest.mark.parametrize("isbloch", list(itertools.product([True, False], repeat=3)))
def test_P2_curl_grad_and_div_curl_vanish(isbloch):
    N = (4, 5, 6)
    M = int(np.prod(N))
    dinv = [crandn(n) for n in N]
    ph = [np.exp(-0.3j * (w + 1)) for w in range(3)]
    for isfwd in (True, False):
        fw = [isfwd] * 3
        C = op.create_curl(fw, dinv, isbloch, ph).to_scipy()
        G = np.zeros((3 * M, M), complex)
        Dv = np.zeros((M, 3 * M), complex)
        for w in range(3):
            D = dense(op.create_d(w, isfwd, N, dinv[w], isbloch[w], ph[w]))
            G[w::3, :] = D
            Dv[:, w::3] = D
        scale = np.abs(C).max() * np.abs(G).max()
        assert np.abs(C @ G).max() <= 1e-13 * scale          # curl grad = 0
        assert np.abs(Dv @ C.toarray()).max() <= 1e-13 * scale  # div curl = 0


# ---- P3 -------------------------------------------------------------------------------
def test_P3_bloch_plane_wave_symbol():
    N = (6, 5, 4)
    d = (0.7, 0.9, 1.1)
    lprim = tuple(np.arange(n + 1) * dd for n, dd in zip(N, d))
    grid = Grid(lprim, (True, True, True))
    kb = np.array([0.31, -0.17, 0.23])
    mvec = np.array([1, 0, -1])
    k = kb + 2 * np.pi * mvec / np.array(grid.L)
    ph = create_e_mikL(kb, grid)
    sdl_e, sdl_m, sei, smi = create_stretched_dls(0.0, grid, ((0, 0, 0), (0, 0, 0)))
    eps = np.zeros(N + (3, 3), complex)
    for v in range(3):
        eps[..., v, v] = 2.5
    omega = 1.3
    Ce, Cm = op.create_curls(sei, smi, (EE,) * 3, grid.isbloch, ph)
    Pe, Pm = op.create_paramops(eps, np.broadcast_to(np.eye(3), N + (3, 3)), sdl_e, sdl_m, sei, smi,
                                (EE,) * 3, grid.isbloch, ph)
    A = op.create_A(EE, omega, Pe, Pm, Ce, Cm)
    # E_v sampled at its Yee location: dual along v, primal along the others
    E0 = np.array([0.3 + 0.1j, -0.7j, 1.1])
    F = np.zeros(N + (3,), complex)
    for v in range(3):
        pos = [grid.l[DUAL if w == v else PRIM][w] for w in range(3)]
        X, Y, Z = np.meshgrid(*pos, indexing="ij")
        F[..., v] = E0[v] * np.exp(-1j * (k[0] * X + k[1] * Y + k[2] * Z))
    y = A.matvec(op.field_arr2vec(F))
    # discrete symbol: forward diff -> (e^{-ik d}-1)/d * e^{+ik d/2}... keep it simple: the
    # Yee-staggered symbols are Dp = (e^{-ikd/2} - e^{+ikd/2})/d for both fwd (E->H) and bwd (H->E)
    D = np.array([(np.exp(-0.5j * kk * dd) - np.exp(0.5j * kk * dd)) / dd for kk, dd in zip(k, d)])
    curl = lambda V: np.cross(D, V)
    Y0 = curl(curl(E0)) - omega ** 2 * 2.5 * E0
    Fy = np.zeros_like(F)
    for v in range(3):
        Fy[..., v] = F[..., v] / E0[v] * Y0[v] if E0[v] != 0 else 0
    assert rel(y, op.field_arr2vec(Fy)) < 1e-13


# ---- P4 -------------------------------------------------------------------------------
@pytest.mark.parametrize("isbloch", [(False, False, False), (True, True, True)])
def test_P4_symmetry_of_A(isbloch):
    N = (4, 5, 3)
    lprim = tuple(np.arange(n + 1) * 1.0 for n in N)
    grid = Grid(lprim, isbloch)
    ph = np.ones(3, complex)
    sdl_e, sdl_m, sei, smi = create_stretched_dls(0.0, grid, ((0, 0, 0), (0, 0, 0)))
    eps = np.zeros(N + (3, 3), complex)
    S = crandn(*N, 3, 3)
    eps[:] = S + S.transpose(0, 1, 2, 4, 3)          # symmetric tensor at every point
    for v in range(3):
        eps[..., v, v] += 4
    mu = np.broadcast_to(np.eye(3), N + (3, 3))
    Ce, Cm = op.create_curls(sei, smi, (EE,) * 3, isbloch, ph)
    Pe, Pm = op.create_paramops(eps, mu, sdl_e, sdl_m, sei, smi, (EE,) * 3, isbloch, ph)
    A = dense(op.create_A(EE, 0.9, Pe, Pm, Ce, Cm))
    if all(isbloch):
        assert np.abs(A - A.T).max() < 1e-13 * np.abs(A).max()
    else:
        # with symmetry BCs the boundary rules ("0's and 2's") break exact symmetry only in rows/cols
        # that touch the first plane; the interior block must be symmetric
        idx = np.arange(np.prod(N)).reshape(N, order="F")
        inner = idx[1:, 1:, 1:].ravel()
        dofs = (3 * inner[:, None] + np.arange(3)[None, :]).ravel()
        B = A[np.ix_(dofs, dofs)]
        assert np.abs(B - B.T).max() < 1e-13 * np.abs(A).max()


def test_P4_symmetrised_with_pml():
    N = (8, 7, 6)
    lprim = tuple(np.arange(n + 1) * 1.0 for n in N)
    grid = Grid(lprim, (True, True, True))
    sdl_e, sdl_m, sei, smi = create_stretched_dls(0.8, grid, ((2, 2, 2), (2, 2, 2)))
    ph = np.ones(3, complex)
    eps = np.zeros(N + (3, 3), complex)
    for v in range(3):
        eps[..., v, v] = 1 + RNG.random(N)
    mu = np.broadcast_to(np.eye(3), N + (3, 3))
    Ce, Cm = op.create_curls(sei, smi, (EE,) * 3, grid.isbloch, ph)
    Pe, Pm = op.create_paramops(eps, mu, sdl_e, sdl_m, sei, smi, (EE,) * 3, grid.isbloch, ph)
    A = dense(op.create_A(EE, 0.8, Pe, Pm, Ce, Cm))
    # length scaling: row of E_v scaled by (edge length along v) * (dual face area) makes A symmetric
    sc = np.zeros(N + (3,), complex)
    for v in range(3):
        a, b = (v + 1) % 3, (v + 2) % 3
        shp = lambda arr, w: np.asarray(arr).reshape([len(arr) if q == w else 1 for q in range(3)])
        sc[..., v] = shp(sdl_m[v], v) * shp(sdl_e[a], a) * shp(sdl_e[b], b)
    S = np.diag(op.field_arr2vec(sc))
    B = S @ A
    assert np.abs(B - B.T).max() < 1e-12 * np.abs(B).max()


# ---- P5 -------------------------------------------------------------------------------
def test_P5_1d_plane_wave_known_answer():
    """Vacuum, x-polarised current sheet J dn at z=0 radiating along +-z into PML (3-D grid that is
    one cell thick and Bloch in x,y).  exp(+iwt): E = -(J dn / 2) * eta0 * exp(-ik|z|), eta0 = 1."""
    Nz, npml = 240, 20
    dz = 0.05
    lam = 1.0
    omega = 2 * np.pi / lam
    lprim = (np.array([0.0, 1.0]), np.array([0.0, 1.0]), (np.arange(Nz + 1) - Nz // 2) * dz)
    grid = Grid(lprim, (True, True, False))
    Npml = ((0, 0, npml), (0, 0, npml))
    sdl_e, sdl_m, sei, smi = create_stretched_dls(omega, grid, Npml)
    ph = np.ones(3, complex)
    N = grid.N
    eps = np.zeros(N + (3, 3), complex)
    mu = np.zeros(N + (3, 3), complex)
    for v in range(3):
        eps[..., v, v] = 1
        mu[..., v, v] = 1
    Ce, Cm = op.create_curls(sei, smi, (EE,) * 3, grid.isbloch, ph)
    Pe, Pm = op.create_paramops(eps, mu, sdl_e, sdl_m, sei, smi, (EE,) * 3, grid.isbloch, ph)
    A = op.create_A(EE, omega, Pe, Pm, Ce, Cm)
    je = create_field_array(N)
    add_src(je, EE, (EE,) * 3, grid, PlaneSrc([0, 0, 1], 0.0, [1, 0, 0]))
    b = op.create_b(EE, omega, Pe, Pm, Ce, Cm, op.field_arr2vec(je), np.zeros(3 * Nz))
    # Ez rows at the z symmetric boundary are all-zero curl rows but keep -w^2 eps: solvable
    e = spla.spsolve(A.to_scipy().tocsc(), b)
    Ex = op.field_vec2arr(e, N)[0, 0, :, 0]
    z = grid.l[PRIM][2]
    inner = slice(npml + 5, Nz - npml - 5)
    # numerical wavenumber of the 2nd-order scheme
    kn = 2 / dz * np.arcsin(omega * dz / 2)
    # discrete Green's function of -d2/dz2 - w^2 with a one-cell sheet K/dz: E0 = -K / (2 cos(kn dz/2))
    ref = -0.5 / np.cos(kn * dz / 2) * np.exp(-1j * kn * np.abs(z))
    err = np.abs(Ex[inner] - ref[inner]).max() / np.abs(ref[inner]).max()
    assert err < 1e-4, err          # only the PML reflection remains
    # sign / time convention: phase must DEcrease away from the source (outgoing wave for exp(+iwt))
    right = np.unwrap(np.angle(Ex[Nz // 2 + 2: Nz - npml - 5]))
    assert np.all(np.diff(right) < 0)


# ---- P6 -------------------------------------------------------------------------------
def _random_problem(N, isbloch, boundft, full_eps, with_mu, ft=EE, npml=1):
    lprim = tuple(np.concatenate(([0.0], np.cumsum(0.5 + RNG.random(n)))) for n in N)
    grid = Grid(lprim, isbloch)
    Npml = (tuple(min(npml, n // 2) for n in N), tuple(min(npml, n // 2) for n in N))
    sdl = create_stretched_dls(0.9 + 0.1j, grid, Npml, boundft)
    kb = np.where(isbloch, RNG.random(3), 0.0)
    ph = create_e_mikL(kb, grid)
    eps = np.zeros(N + (3, 3), complex)
    mu = np.zeros(N + (3, 3), complex)
    for v in range(3):
        eps[..., v, v] = 2 + crandn(*N) * 0.3
        mu[..., v, v] = (1.5 + crandn(*N) * 0.2) if with_mu else 1.0
    if full_eps:
        for v in range(3):
            for u in range(3):
                if u != v:
                    eps[..., v, u] = crandn(*N) * 0.3
    return grid, sdl, ph, eps, mu


CASES = []
for _N in [(1, 1, 1), (2, 1, 3), (3, 2, 1), (5, 3, 2), (3, 5, 8)]:
    for _bl in itertools.product([True, False], repeat=3):
        CASES.append((_N, _bl))


@pytest.mark.parametrize("N,isbloch", CASES)
def test_P6_csc_vs_matrix_free(N, isbloch):
    for boundft in itertools.product([EE, HH], repeat=3):
        for full_eps, with_mu, ft in ((False, False, EE), (True, True, EE), (False, True, HH)):
            grid, (sdl_e, sdl_m, sei, smi), ph, eps, mu = _random_problem(N, isbloch, boundft, full_eps, with_mu)
            omega = 1.1 - 0.05j
            Ce, Cm = op.create_curls(sei, smi, boundft, isbloch, ph)
            Pe, Pm = op.create_paramops(eps, mu, sdl_e, sdl_m, sei, smi, boundft, isbloch, ph)
            A = op.create_A(ft, omega, Pe, Pm, Ce, Cm)
            mf = MatFreeOperator(ft, omega, eps, mu, sdl_e, sdl_m, boundft, isbloch, ph)
            x = crandn(3 * int(np.prod(N)))
            y1, y2 = A.matvec(x), mf(x)
            assert rel(y2, y1) < 1e-13, (N, isbloch, boundft, full_eps, ft)


def test_P6_soa_ordering_and_weighted_out():
    N, isbloch, boundft = (4, 3, 5), (True, False, True), (EE, HH, EE)
    grid, (sdl_e, sdl_m, sei, smi), ph, eps, mu = _random_problem(N, isbloch, boundft, True, True)
    for cmpfirst in (True, False):
        for wo in (False, True):
            Ce, Cm = op.create_curls(sei, smi, boundft, isbloch, ph, cmpfirst)
            Pe, Pm = op.create_paramops(eps, mu, sdl_e, sdl_m, sei, smi, boundft, isbloch, ph, cmpfirst, wo)
            A = op.create_A(EE, 0.7, Pe, Pm, Ce, Cm)
            mf = MatFreeOperator(EE, 0.7, eps, mu, sdl_e, sdl_m, boundft, isbloch, ph, cmpfirst, wo)
            x = crandn(3 * int(np.prod(N)))
            assert rel(mf(x), A.matvec(x)) < 1e-13


# ---- pattern rules -----------------------------------------------------------
def test_pattern_rules_interior_13_and_zero_dropping():
    N = (5, 6, 7)
    lprim = tuple(np.arange(n + 1) * 1.0 for n in N)
    eps = np.zeros(N + (3, 3), complex)
    for v in range(3):
        eps[..., v, v] = 2.0
    mu = np.broadcast_to(np.eye(3), N + (3, 3))
    for isbloch in ((True, True, True), (False, False, False)):
        grid = Grid(lprim, isbloch)
        sdl_e, sdl_m, sei, smi = create_stretched_dls(0.0, grid, ((0, 0, 0), (0, 0, 0)))
        ph = np.ones(3, complex)
        Ce, Cm = op.create_curls(sei, smi, (EE,) * 3, isbloch, ph)
        Pe, Pm = op.create_paramops(eps, mu, sdl_e, sdl_m, sei, smi, (EE,) * 3, isbloch, ph)
        A0 = op.create_A(EE, 0.0, Pe, Pm, Ce, Cm)       # w == 0: structural pattern, zeros kept
        A1 = op.create_A(EE, 1.0, Pe, Pm, Ce, Cm)       # w != 0: exact zeros dropped
        assert np.all(np.diff(A0.colptr) == 13)          # periodic structural pattern everywhere
        assert np.all(np.diff(Ce.colptr) == 4) and np.all(np.diff(Cm.colptr) == 4)
        if all(isbloch):
            assert A1.nnz == A0.nnz
        else:
            assert A1.nnz < A0.nnz and np.all(A1.nzval != 0)
            assert (A0.nzval == 0).any()
        cp, rv = A1.julia_pattern()
        assert cp[0] == 1 and cp.dtype == np.int64 and rv.min() >= 1
        for j in range(0, A1.shape[1], 17):
            rows = A1.rowval[A1.colptr[j]:A1.colptr[j + 1]]
            assert np.all(np.diff(rows) > 0)


# ---- P7 -------------------------------------------------------------------------------
def test_P7_dipole_in_pml_box_direct_solve():
    n, npml = 22, 6
    d = 1.0
    lam = 10.0 * d
    omega = 2 * np.pi / lam
    lp = (np.arange(n + 1) - n / 2) * d
    grid = Grid((lp, lp, lp), (False, False, False))
    Npml = ((npml,) * 3, (npml,) * 3)
    sdl_e, sdl_m, sei, smi = create_stretched_dls(omega, grid, Npml)
    N = grid.N
    eps = np.zeros(N + (3, 3), complex)
    mu = np.zeros(N + (3, 3), complex)
    for v in range(3):
        eps[..., v, v] = 1
        mu[..., v, v] = 1
    ph = np.ones(3, complex)
    Ce, Cm = op.create_curls(sei, smi, (EE,) * 3, grid.isbloch, ph)
    Pe, Pm = op.create_paramops(eps, mu, sdl_e, sdl_m, sei, smi, (EE,) * 3, grid.isbloch, ph)
    A = op.create_A(EE, omega, Pe, Pm, Ce, Cm)
    je = create_field_array(N)
    add_src(je, EE, (EE,) * 3, grid, PointSrc([0.0, 0.0, 0.0], [0, 0, 1]))
    b = op.create_b(EE, omega, Pe, Pm, Ce, Cm, op.field_arr2vec(je), np.zeros(3 * n ** 3))
    As = A.to_scipy().tocsc()
    e = spla.splu(As).solve(b)
    assert rel(As @ e, b) < 1e-10
    Ez = np.abs(op.field_vec2arr(e, N)[..., 2])
    c = n // 2
    # field decays into the PML: amplitude at the outer wall << amplitude at the PML entrance
    assert Ez[c, c - 1, 1] < 0.05 * Ez[c, c - 1, npml]
    assert Ez[1, c - 1, c] < 0.05 * Ez[npml, c - 1, c]


# ---- P8 -------------------------------------------------------------------------------
@pytest.mark.parametrize("boundft", [(EE, EE, EE), (HH, HH, HH), (EE, HH, EE)])
def test_P8_cavity_spectrum_pins_the_symmetry_boundaries(boundft):
    """Closed box with symmetry boundaries on all six faces, uniform grid, no PML, w = 0: A = Cm Ce is the discrete
    curl-curl of a cavity.  Whatever the wall type per axis (boundft = EE: electric wall, HH: magnetic wall), the 1-D
    Laplacians D- D+ / D+ D- have eigenvalues k_w(m)^2 = (2/d_w sin(m pi / 2N_w))^2, m = 0..N_w-1, so every non-zero
    eigenvalue of A must be a sum kx^2 + ky^2 + kz^2 of such values, and every sum with all three mode numbers
    >= 1 must occur at least twice (two polarisations).  (Sums with two zero mode numbers occur as well: the
    tangential unknowns ON a wall stay in the system - boundary rules zero values, never positions - and couple
    among themselves within the wall plane.)  This pins what the boundary rows of create_d do at both ends, which
    the adjointness / nilpotency identities P1-P2 leave open."""
    N, d = (3, 4, 5), (1.0, 0.8, 1.25)
    lprim = tuple(np.arange(n + 1) * dw for n, dw in zip(N, d))
    grid = Grid(lprim, (False, False, False))
    sdl_e, sdl_m, sei, smi = create_stretched_dls(0.0, grid, ((0, 0, 0), (0, 0, 0)), boundft)
    ph = np.ones(3, complex)
    Ce, Cm = op.create_curls(sei, smi, boundft, grid.isbloch, ph)
    A = (Cm.to_scipy() @ Ce.to_scipy()).toarray()
    lam = np.linalg.eigvals(A)
    assert np.abs(lam.imag).max() < 1e-10                      # real spectrum
    lam = np.sort(lam.real)
    assert lam.min() > -1e-10                                  # positive semi-definite
    nz = lam[lam > 1e-8]
    k2 = [np.array([(2 / dw * np.sin(m * np.pi / (2 * n))) ** 2 for m in range(n)]) for n, dw in zip(N, d)]
    allowed = np.array([k2[0][a] + k2[1][b] + k2[2][c] for a in range(N[0]) for b in range(N[1]) for c in range(N[2])
                        if a + b + c > 0])
    dist = np.abs(nz[:, None] - allowed[None, :]).min(axis=1)
    assert dist.max() < 1e-9, dist.max()                       # nothing outside the separable spectrum
    full = np.array([k2[0][a] + k2[1][b] + k2[2][c] for a in range(1, N[0]) for b in range(1, N[1])
                     for c in range(1, N[2])])
    for v in full:
        assert np.sum(np.abs(nz - v) < 1e-9) >= 2, v           # both polarisations of every volume mode
    # the null space is exactly the discrete gradients + the decoupled wall unknowns: rank = #non-zero eigenvalues
    assert np.linalg.matrix_rank(A, tol=1e-9) == nz.size


# ---- P9 -------------------------------------------------------------------------------
def test_P9_paramop_preserves_uniform_fields_on_nonuniform_grids():
    """Default arrangement (boundft all-EE, unweighted output mean), non-uniform periodic grid (k = 0), one constant
    full 3x3 tensor everywhere, uniform field E0: the two averaging steps of create_paramop must reproduce eps * E0
    exactly.  The input mean takes E_w from the dual points i-1, i to the primal corner i with weights
    dl_dual[i-1], dl_dual[i] over 2 dl_prim[i]; these sum to one only because a primal cell is made of the two
    adjacent half dual cells - the identity fails if a cell size is paired with the wrong neighbour or taken from the
    wrong (primal/dual) array.  (The mirror-image statement for primal -> dual means is NOT an identity on a
    non-uniform grid - primal points are not midpoints of dual points - so it is not asserted for boundft = HH or for
    the weighted output mean.)"""
    N = (4, 5, 3)
    rng = np.random.default_rng(9)
    lprim = tuple(np.concatenate(([0.0], np.cumsum(0.5 + rng.random(n)))) for n in N)
    grid = Grid(lprim, (True, True, True))
    boundft = (EE, EE, EE)
    sdl_e, sdl_m, sei, smi = create_stretched_dls(0.0, grid, ((0, 0, 0), (0, 0, 0)), boundft)
    T = rng.standard_normal((3, 3)) + 1j * rng.standard_normal((3, 3))
    eps = np.broadcast_to(T, N + (3, 3)).copy()
    mu = np.broadcast_to(np.eye(3), N + (3, 3))
    ph = np.ones(3, complex)
    Pe, _ = op.create_paramops(eps, mu, sdl_e, sdl_m, sei, smi, boundft, grid.isbloch, ph, True, False)
    E0 = np.array([0.7 - 0.2j, -1.1 + 0.4j, 0.3 + 0.9j])
    e = op.field_arr2vec(np.broadcast_to(E0, N + (3,)).copy())
    want = op.field_arr2vec(np.broadcast_to(T @ E0, N + (3,)).copy())
    assert rel(Pe.matvec(e), want) < 1e-13
    # the same inputs through the matrix-free oracle
    mf = MatFreeOperator(EE, 1.0, eps, None, sdl_e, sdl_m, boundft, grid.isbloch, ph)
    assert rel(mf.arr2vec(mf.Peps(mf.vec2arr(e))), want) < 1e-13
