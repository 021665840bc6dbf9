"""Material-parameter pipeline - oracle restatement, TEST INFRASTRUCTURE ONLY.

Reference: `calc_matparams!(mdl::ModelFull)` (src/model/full.jl:16-70) = `assign_param!` + `smooth_param!` of the
un-vendored MaxwellBase ^0.1.6 with shapes from GeometryPrimitives (Project.toml:10,15; neither is on disk).
PARITY UNPINNED: the reference holds no test or fixture for this path and the code that does the arithmetic is not in
the tree, so this file restates the published algorithm (Kottke, Farjadpour, Johnson, PRE 77, 036611: subpixel
smoothing of anisotropic interfaces) in the structure the call site shows:

  * where values live (full.jl:58-67, model.jl:141-158): the diagonal entry P_vv at the Yee location of field component
    v, all off-diagonal entries at the location whose grid type is that of the field PLANES on every axis (the voxel
    corners of E for eps: the kernels' "corner-located" off-diagonals);
  * the voxel of a location is bounded, per axis, by the neighbouring points of the other grid type; its 8 corners
    are therefore the locations of the other field (which is why full.jl:58-61 fills the object-index arrays of mu
    while assigning eps and vice versa);
  * per voxel: object at each corner = the LAST added shape containing it (ghost points map back into the domain:
    periodic wrap for Bloch, mirror for symmetry boundaries); one material in the voxel -> that material; two
    materials -> Kottke average `tau^-1(<tau(P)>)` in the frame of the interface normal, with the normal and the
    foreground volume fraction from the foreground shape's nearest surface point to the voxel centre when exactly two
    objects meet (plane-cut volume, exact), else from the corner occupancy; three or more -> harmonic mean over the
    corners (arithmetic mean when the field is orthogonal to the shape dimensions, model.jl:65-69).

Choices the in-tree evidence cannot fix (tie-breaking on shape surfaces, nearest-point rules at box edges, the
degenerate-normal threshold of the volume fraction) are documented where they are made; the CUDA path
(csrc/matparams.cu) makes the same ones.
"""
from __future__ import annotations

import itertools

import numpy as np

from .grid import EE, HH, PRIM, DUAL, Grid, ft2gt

BOX, BALL, CYL = 0, 1, 2
VOLFRAC_TOL = 1e-6     # an axis with |n_w| h_w below this fraction of the largest one is treated as parallel to the plane


class Box:
    """Axis-aligned cuboid: centre c, half-widths r."""
    kind = BOX

    def __init__(self, c, r):
        self.c = np.asarray(c, float)
        self.r = np.asarray(r, float)
        self.axis = 0


class Ball:
    kind = BALL

    def __init__(self, c, radius):
        self.c = np.asarray(c, float)
        self.r = np.array([radius, 0.0, 0.0])
        self.axis = 0


class Cylinder:
    """Right circular cylinder along a coordinate axis: centre c, radius, half-height h."""
    kind = CYL

    def __init__(self, c, radius, h, axis=2):
        self.c = np.asarray(c, float)
        self.r = np.array([radius, h, 0.0])
        self.axis = int(axis)


def contains(s, x):
    """closed sets: points on the surface belong to the shape.  x: (..., 3)"""
    d = x - s.c
    if s.kind == BOX:
        return np.all(np.abs(d) <= s.r, axis=-1)
    if s.kind == BALL:
        return d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2] <= s.r[0] * s.r[0]
    a = s.axis
    b, c = (a + 1) % 3, (a + 2) % 3
    return (np.abs(d[..., a]) <= s.r[1]) & (d[..., b] * d[..., b] + d[..., c] * d[..., c] <= s.r[0] * s.r[0])


def surfpt_nearby(s, x0):
    """(r0, nout): the point of the shape's surface nearest to x0 and the outward unit normal there (for x0 outside:
    the direction from r0 to x0, which is the normal on smooth parts and the natural choice at edges)."""
    d = x0 - s.c
    if s.kind == BALL:
        nd = np.sqrt(d @ d)
        n = d / nd if nd > 0 else np.array([1.0, 0.0, 0.0])
        return s.c + s.r[0] * n, n
    if s.kind == BOX:
        q = np.abs(d) - s.r
        if np.all(q <= 0):
            a = int(np.argmax(q))                       # first of equally near faces
            n = np.zeros(3)
            n[a] = 1.0 if d[a] >= 0 else -1.0
            r0 = x0.copy()
            r0[a] = s.c[a] + n[a] * s.r[a]
            return r0, n
        r0 = s.c + np.clip(d, -s.r, s.r)
        v = x0 - r0
        return r0, v / np.sqrt(v @ v)
    a = s.axis
    b, c = (a + 1) % 3, (a + 2) % 3
    R, h = s.r[0], s.r[1]
    rho = np.sqrt(d[b] * d[b] + d[c] * d[c])
    qa, qr = abs(d[a]) - h, rho - R
    if qa <= 0 and qr <= 0:
        n = np.zeros(3)
        r0 = x0.copy()
        if qa > qr:                                      # cap nearer than the side
            n[a] = 1.0 if d[a] >= 0 else -1.0
            r0[a] = s.c[a] + n[a] * h
        else:
            if rho > 0:
                n[b], n[c] = d[b] / rho, d[c] / rho
            else:
                n[b] = 1.0
            r0[b], r0[c] = s.c[b] + R * n[b], s.c[c] + R * n[c]
        return r0, n
    r0 = x0.copy()
    r0[a] = s.c[a] + min(max(d[a], -h), h)
    f = 1.0 if rho <= R else R / rho
    r0[b], r0[c] = s.c[b] + f * d[b], s.c[c] + f * d[c]
    v = x0 - r0
    return r0, v / np.sqrt(v @ v)


def volfrac(lo, hi, n, r0):
    """Fraction of the box [lo,hi] on the inner side of the plane through r0 with outward normal n (n.(x-r0) <= 0):
    exact polynomial formula (inclusion-exclusion over the box corners), dimension reduced for axes (nearly) parallel
    to the plane."""
    h = hi - lo
    a = np.abs(n) * h
    dmin = sum(min(n[w] * lo[w], n[w] * hi[w]) for w in range(3))
    d = float(n @ r0) - dmin                              # plane offset measured from the innermost corner
    amax = a.max()
    act = [w for w in range(3) if a[w] > VOLFRAC_TOL * amax]
    k = len(act)
    tot = 0.0
    for m in range(k + 1):
        for S in itertools.combinations(act, m):
            t = d - sum(a[w] for w in S)
            if t > 0:
                tot += (-1.0) ** m * t ** k
    f = tot / (np.prod([a[w] for w in act]) * (1, 1, 2, 6)[k])
    return min(max(f, 0.0), 1.0)


def _frame(n):
    """orthonormal S = [n t1 t2] (columns); t1 from the coordinate axis least aligned with n"""
    e = np.zeros(3)
    e[int(np.argmin(np.abs(n)))] = 1.0
    t1 = e - (e @ n) * n
    t1 /= np.sqrt(t1 @ t1)
    t2 = np.cross(n, t1)
    return np.stack([n, t1, t2], axis=1)


def _tau(P):
    T = np.empty((3, 3), complex)
    T[0, 0] = -1.0 / P[0, 0]
    T[0, 1:] = P[0, 1:] / P[0, 0]
    T[1:, 0] = P[1:, 0] / P[0, 0]
    T[1:, 1:] = P[1:, 1:] - np.outer(P[1:, 0], P[0, 1:]) / P[0, 0]
    return T


def _tau_inv(T):
    P = np.empty((3, 3), complex)
    P[0, 0] = -1.0 / T[0, 0]
    P[0, 1:] = -T[0, 1:] / T[0, 0]
    P[1:, 0] = -T[1:, 0] / T[0, 0]
    P[1:, 1:] = T[1:, 1:] - np.outer(T[1:, 0], T[0, 1:]) / T[0, 0]
    return P


def kottke_avg_param(P1, P2, n12, rvol1):
    """Kottke's average of P1 (volume fraction rvol1) and P2 across an interface with unit normal n12."""
    S = _frame(np.asarray(n12, float))
    T = rvol1 * _tau(S.T @ P1 @ S) + (1.0 - rvol1) * _tau(S.T @ P2 @ S)
    return S @ _tau_inv(T) @ S.T


def _ghosted(grid: Grid):
    """per axis: (ghosted primal [N+1], ghosted dual [N+1])"""
    return [(grid.lprim_g[w], grid.ldual_g[w]) for w in range(3)]


def _to_domain(grid: Grid, x):
    """tau-transform of a (ghost) point: wrap for Bloch axes, mirror at symmetry boundaries"""
    y = np.array(x, float)
    for w in range(3):
        lo, hi = grid.bounds[0][w], grid.bounds[1][w]
        if y[w] < lo:
            y[w] = y[w] + grid.L[w] if grid.isbloch[w] else 2.0 * lo - y[w]
        elif y[w] > hi:
            y[w] = y[w] - grid.L[w] if grid.isbloch[w] else 2.0 * hi - y[w]
    return y


def location_gt(ft, boundft, v):
    """grid type per axis of the location of entry (v,v) (v = 0..2) or of the off-diagonal entries (v = 3)"""
    g = [ft2gt(ft, boundft[w]) for w in range(3)]
    if v < 3:
        g[v] = DUAL - g[v]
    return g


def calc_matparams(grid: Grid, boundft, ft, shapes, pinds, params, field_ortho_shape=False):
    """Returns arr[i,j,k,v,u] (Nx,Ny,Nz,3,3): the smoothed parameter array of field type ft (EE: eps, HH: mu).
    shapes: ordered list (later shapes lie on top); pinds[o]: parameter index of shape o; params[p]: 3x3 tensors.
    Every voxel corner must lie in at least one shape (add a background Box first, as reference users do)."""
    N = grid.N
    gl = _ghosted(grid)
    params = [np.asarray(P, complex).reshape(3, 3) for P in params]
    symmetric = all(np.array_equal(P, P.T) for P in params)   # then the average is symmetric: upper triangle to both places
    out = np.zeros(N + (3, 3), complex)
    for v in range(4):
        gts = location_gt(ft, boundft, v)
        # voxel bounds per axis: PRIM location i -> [dual_g[i], dual_g[i+1]]; DUAL location i -> [prim_g[i], prim_g[i+1]]
        edges = [gl[w][1] if gts[w] == PRIM else gl[w][0] for w in range(3)]
        for i, j, k in itertools.product(range(N[0]), range(N[1]), range(N[2])):
            idx = (i, j, k)
            lo = np.array([edges[w][idx[w]] for w in range(3)])
            hi = np.array([edges[w][idx[w] + 1] for w in range(3)])
            P = _voxel_param(grid, lo, hi, shapes, pinds, params, field_ortho_shape)
            if v < 3:
                out[i, j, k, v, v] = P[v, v]
            else:
                for a, b in itertools.permutations(range(3), 2):
                    out[i, j, k, a, b] = P[b, a] if (symmetric and a > b) else P[a, b]
    return out


def _object_at(grid, shapes, x):
    xt = _to_domain(grid, x)
    for o in range(len(shapes) - 1, -1, -1):
        if contains(shapes[o], xt):
            return o
    raise ValueError(f"no shape covers the point {x}: add a background shape first")


def _voxel_param(grid, lo, hi, shapes, pinds, params, field_ortho_shape):
    corners = [np.array([(lo, hi)[(c >> w) & 1][w] for w in range(3)]) for c in range(8)]
    oc = [_object_at(grid, shapes, x) for x in corners]
    pc = [pinds[o] for o in oc]
    distinct = sorted(set(pc))
    if len(distinct) == 1:
        return params[pc[0]]
    if len(distinct) >= 3:
        if field_ortho_shape:
            return sum(params[p] for p in pc) / 8.0
        return np.linalg.inv(sum(np.linalg.inv(params[p]) for p in pc) / 8.0)
    o_fg = max(oc)                                          # the topmost object is the foreground
    p_fg = pinds[o_fg]
    p_bg = distinct[0] if distinct[1] == p_fg else distinct[1]
    x0 = 0.5 * (lo + hi)
    if len(set(oc)) == 2:
        r0, nout = surfpt_nearby(shapes[o_fg], _to_domain(grid, x0))
        rvol = volfrac(lo, hi, nout, r0 + (x0 - _to_domain(grid, x0)))
    else:
        # two materials but more than two objects: normal and volume fraction from the corner occupancy
        fg = [p == p_fg for p in pc]
        rvol = sum(fg) / 8.0
        nvec = np.zeros(3)
        for c in range(8):
            sgn = -1.0 if fg[c] else 1.0
            nvec += sgn * np.array([1.0 if (c >> w) & 1 else -1.0 for w in range(3)])
        nn = np.sqrt(nvec @ nvec)
        if nn == 0.0:
            if field_ortho_shape:
                return sum(params[p] for p in pc) / 8.0
            return np.linalg.inv(sum(np.linalg.inv(params[p]) for p in pc) / 8.0)
        nout = nvec / nn
    if field_ortho_shape:
        return rvol * params[p_fg] + (1.0 - rvol) * params[p_bg]
    return kottke_avg_param(params[p_fg], params[p_bg], nout, rvol)
