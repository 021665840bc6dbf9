"""Oracle of the material pipeline (oracle/matparams.py) - PARITY UNPINNED: the reference holds no
test or fixture for calc_matparams! (full.jl:16-70) and the arithmetic lives in un-vendored packages, so the
restatement is pinned by what the published algorithm (Kottke et al., PRE 77, 036611) implies:

  M1  the plane-cut volume fraction is exact (closed forms for axis-aligned and 45-degree cuts, Monte Carlo otherwise);
  M2  Kottke's average of two isotropic media is arithmetic along the interface and harmonic across it, for any normal;
      it is symmetric in (P1, f) <-> (P2, 1-f) with the normal reversed, and returns P for P1 == P2;
  M3  a slab cut by an axis-aligned plane gives the textbook per-component fill-fraction means at the right Yee
      locations (this also pins WHICH voxel belongs to which entry), no off-diagonal entries;
  M4  a sphere gives off-diagonal entries only on its surface voxels, with the symmetry of the sphere;
  M5  voxels between two objects of the same material are not smoothed; later objects lie on top;
  M6  uncovered points are an error.
"""
import itertools

import numpy as np
import pytest

from oracle import matparams as mp
from oracle.grid import Grid, EE, HH


def test_M1_volume_fraction():
    lo, hi = np.array([0.0, 0.0, 0.0]), np.array([1.0, 2.0, 0.5])
    # axis-aligned cut
    assert mp.volfrac(lo, hi, np.array([0.0, 1.0, 0.0]), np.array([9.0, 0.5, 9.0])) == pytest.approx(0.25, abs=1e-15)
    assert mp.volfrac(lo, hi, np.array([0.0, -1.0, 0.0]), np.array([9.0, 0.5, 9.0])) == pytest.approx(0.75, abs=1e-15)
    # 45-degree cut through the centre of a unit square cross-section
    lo2, hi2 = np.zeros(3), np.ones(3)
    n = np.array([1.0, 1.0, 0.0]) / np.sqrt(2)
    assert mp.volfrac(lo2, hi2, n, np.array([0.5, 0.5, 0.3])) == pytest.approx(0.5, abs=1e-14)
    # corner tetrahedron: x + y + z <= 0.3 in the unit cube -> 0.3^3 / 6
    n = np.ones(3) / np.sqrt(3)
    assert mp.volfrac(lo2, hi2, n, np.array([0.3, 0.0, 0.0])) == pytest.approx(0.3 ** 3 / 6, rel=1e-12)
    # complement rule and Monte Carlo for generic planes
    rng = np.random.default_rng(1)
    for _ in range(6):
        lo = rng.random(3)
        hi = lo + 0.3 + rng.random(3)
        n = rng.standard_normal(3)
        n /= np.linalg.norm(n)
        r0 = lo + (hi - lo) * rng.random(3)
        f = mp.volfrac(lo, hi, n, r0)
        assert f + mp.volfrac(lo, hi, -n, r0) == pytest.approx(1.0, abs=1e-12)
        pts = lo + (hi - lo) * rng.random((200000, 3))
        assert f == pytest.approx(np.mean((pts - r0) @ n <= 0), abs=5e-3)
    # planes that miss the box
    assert mp.volfrac(lo2, hi2, np.array([1.0, 0, 0]), np.array([-1.0, 0, 0])) == 0.0
    assert mp.volfrac(lo2, hi2, np.array([1.0, 0, 0]), np.array([2.0, 0, 0])) == 1.0


def test_M2_kottke_average():
    rng = np.random.default_rng(2)
    for _ in range(5):
        e1, e2, f = 1 + 11 * rng.random(), 1 + 3 * rng.random(), rng.random()
        n = rng.standard_normal(3)
        n /= np.linalg.norm(n)
        K = mp.kottke_avg_param(e1 * np.eye(3), e2 * np.eye(3), n, f)
        ar, hm = f * e1 + (1 - f) * e2, 1 / (f / e1 + (1 - f) / e2)
        assert np.abs(K - (ar * (np.eye(3) - np.outer(n, n)) + hm * np.outer(n, n))).max() < 1e-13
        P1 = np.diag(1 + rng.random(3)) + 0.2 * (rng.standard_normal((3, 3)) + 1j * rng.standard_normal((3, 3)))
        P2 = np.diag(3 + rng.random(3)) + 0.2 * (rng.standard_normal((3, 3)) + 1j * rng.standard_normal((3, 3)))
        A = mp.kottke_avg_param(P1, P2, n, f)
        assert np.abs(A - mp.kottke_avg_param(P2, P1, -n, 1 - f)).max() < 1e-13
        assert np.abs(mp.kottke_avg_param(P1, P1, n, f) - P1).max() < 1e-13
        assert np.abs(mp.kottke_avg_param(P1, P2, n, 1.0) - P1).max() < 1e-13
        # the tangential field components and the normal flux component are continuous: for E in the interface
        # plane the average is the plain arithmetic one
        t = np.cross(n, rng.standard_normal(3))
        t /= np.linalg.norm(t)
        Pi1, Pi2 = f * P1 + (1 - f) * P2, A
        if np.allclose(P1, np.diag(np.diag(P1))):     # only exact for media without normal-tangential coupling
            assert abs(t @ Pi1 @ t - t @ Pi2 @ t) < 1e-12


def _slab_model(isbloch=(True, True, False), boundft=(EE, EE, EE)):
    lp = np.arange(9.0)
    g = Grid((lp, lp, lp), isbloch)
    shapes = [mp.Box([4, 4, 4], [10, 10, 10]), mp.Box([4, 4, 1.3], [10, 10, 2.0])]   # eps = 4 for z <= 3.3
    return g, shapes


def test_M3_planar_interface_gives_fill_fraction_means_at_the_yee_locations():
    g, shapes = _slab_model()
    arr = mp.calc_matparams(g, (EE, EE, EE), EE, shapes, [0, 1], [np.eye(3), 4 * np.eye(3)])
    # E_x, E_y sit on the primal z planes k: voxel [k-1/2, k+1/2]; plane 3 is filled to 0.8 -> arithmetic mean
    for v in (0, 1):
        assert np.allclose(arr[2, 5, :, v, v], [4, 4, 4, 0.8 * 4 + 0.2, 1, 1, 1, 1], atol=1e-13)
    # E_z sits at k+1/2: voxel [k, k+1]; voxel 3 is filled to 0.3 -> harmonic mean (normal component)
    assert np.allclose(arr[2, 5, :, 2, 2], [4, 4, 4, 1 / (0.3 / 4 + 0.7), 1, 1, 1, 1], atol=1e-13)
    assert not arr[..., 0, 1].any() and not arr[..., 2, 0].any()
    # mu locations are the dual ones: H_z on the primal z planes ... swap of the two patterns
    arr_m = mp.calc_matparams(g, (EE, EE, EE), HH, shapes, [0, 1], [np.eye(3), 4 * np.eye(3)])
    assert np.allclose(arr_m[2, 5, :, 2, 2], [4, 4, 4, 1 / (0.8 / 4 + 0.2), 1, 1, 1, 1], atol=1e-13)
    assert np.allclose(arr_m[2, 5, :, 0, 0], [4, 4, 4, 0.3 * 4 + 0.7, 1, 1, 1, 1], atol=1e-13)
    # boundft = HH on z moves the E planes to the dual points
    arr_h = mp.calc_matparams(g, (EE, EE, HH), EE, shapes, [0, 1], [np.eye(3), 4 * np.eye(3)])
    assert np.allclose(arr_h[2, 5, :, 0, 0], arr_m[2, 5, :, 0, 0], atol=1e-13)


def test_M4_sphere_symmetry_and_offdiagonal_support():
    lp = np.arange(13.0) - 6.0
    g = Grid((lp, lp, lp), (False, False, False))
    shapes = [mp.Box([0, 0, 0], [20, 20, 20]), mp.Ball([0, 0, 0], 3.3)]
    arr = mp.calc_matparams(g, (EE, EE, EE), EE, shapes, [0, 1], [np.eye(3), 9 * np.eye(3)])
    # corner-located off-diagonal entries: node (i,j,k) at lp; mirror symmetry x -> -x maps node index i -> 12 - i
    exy = arr[..., 0, 1]
    assert np.abs(exy).max() > 0.1
    assert np.allclose(exy[1:, 1:, 1:], -exy[1:, 1:, 1:][::-1, :, :], atol=1e-12)     # odd under x -> -x
    assert np.allclose(exy[1:, 1:, 1:], exy[1:, 1:, 1:][:, :, ::-1], atol=1e-12)      # even under z -> -z
    assert np.allclose(arr[..., 0, 1], arr[..., 1, 0], atol=1e-13)                    # symmetric tensor stays symmetric
    # x <-> y exchange symmetry of the sphere
    assert np.allclose(arr[..., 0, 0], arr[..., 1, 1].transpose(1, 0, 2), atol=1e-12)
    # off-diagonal entries only where the surface passes: |r| within one cell diagonal of the radius
    X, Y, Z = np.meshgrid(lp[:-1], lp[:-1], lp[:-1], indexing="ij")
    R = np.sqrt(X ** 2 + Y ** 2 + Z ** 2)
    assert not exy[np.abs(R - 3.3) > np.sqrt(3) / 2 + 1e-9].any()
    # deep inside / far outside: the plain materials
    assert np.allclose(arr[6, 6, 6], 9 * np.eye(3)) and np.allclose(arr[0, 0, 0], np.eye(3))


def test_M5_same_material_objects_and_stacking_order():
    lp = np.arange(9.0)
    g = Grid((lp, lp, lp), (True, True, True))
    bg = mp.Box([4, 4, 4], [10, 10, 10])
    a, b = mp.Box([2.6, 4, 4], [2.6, 10, 10]), mp.Box([5.0, 4, 4], [1.6, 10, 10])   # overlap over x in [3.4, 5.2]
    P = [np.eye(3), 5 * np.eye(3), 2 * np.eye(3)]
    same = mp.calc_matparams(g, (EE,) * 3, EE, [bg, a, b], [0, 1, 1], P)
    one = mp.calc_matparams(g, (EE,) * 3, EE, [bg, mp.Box([3.3, 4, 4], [3.3, 10, 10])], [0, 1], P)
    assert np.allclose(same, one, atol=1e-13)                     # the seam between a and b is invisible
    ab = mp.calc_matparams(g, (EE,) * 3, EE, [bg, a, b], [0, 1, 2], P)
    ba = mp.calc_matparams(g, (EE,) * 3, EE, [bg, b, a], [0, 2, 1], P)
    assert ab[4, 3, 3, 1, 1] == 2.0 and ba[4, 3, 3, 1, 1] == 5.0  # the voxel of x = 4 lies in the overlap: the later object wins


def test_M6_uncovered_point_is_an_error():
    lp = np.arange(5.0)
    g = Grid((lp, lp, lp), (False, False, False))
    with pytest.raises(ValueError):
        mp.calc_matparams(g, (EE,) * 3, EE, [mp.Ball([2, 2, 2], 1.0)], [0], [np.eye(3)])
