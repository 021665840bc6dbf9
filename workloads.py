"""Synthetic inputs for the BASELINE.json configurations (SURVEY.md §8d recipes) - bench/test infrastructure.

Everything here is host-side numpy and uses only the PRODUCT's host modules (grid geometry, PML), never
the oracle.  Lengths are in nm; lambda = 1550 nm; boundft = (EE,EE,EE); order_cmpfirst = True; mu = 1.

Each builder returns a dict:
    N (global), isbloch, kbloch, Npml, grid, omega, sdl_e, sdl_m, e_mikL, eps (this rank's slab
    [Nx,Ny,k1-k0,3,3]), k0, k1, full_eps, name
The reference has no mode solver and no TF/SF source (model.jl:30-31, README.md:40); sources are the
stand-ins of §8d and are built on demand by `rhs()` from PlaneSrc / PointSrc.
"""
import numpy as np

import maxwellfdm_jl_b200 as fb

LAMBDA = 1550.0
OMEGA = 2 * np.pi / LAMBDA


def _common(N, delta, isbloch, Npml, kbloch=(0.0, 0.0, 0.0)):
    lprim = tuple((np.arange(n + 1) - n / 2.0) * delta for n in N)
    grid = fb.Grid(lprim, isbloch)
    mdl = fb.ModelFull(grid)
    fb.set_wpml(mdl, OMEGA)
    fb.set_Npml(mdl, Npml)
    fb.set_kbloch(mdl, kbloch)
    sdl_e, sdl_m, _, _ = fb.create_stretched_dls(mdl)
    return dict(N=tuple(N), isbloch=tuple(isbloch), kbloch=tuple(kbloch), Npml=Npml, grid=grid, omega=OMEGA,
                sdl_e=sdl_e, sdl_m=sdl_m, e_mikL=fb.create_e_mikL(mdl), model=mdl)


def _fill_1d(centers, delta, half_width):
    """fraction of the cell [c-d/2, c+d/2] that lies inside |u| <= half_width"""
    lo = np.maximum(centers - delta / 2, -half_width)
    hi = np.minimum(centers + delta / 2, half_width)
    return np.clip(hi - lo, 0.0, None) / delta


def _mix(fill, e_in, e_out, harmonic):
    a = fill * e_in + (1 - fill) * e_out
    h = 1.0 / (fill / e_in + (1 - fill) / e_out)
    return np.where(harmonic, h, a)


def c1_vacuum_box(N=(40, 40, 40), k0=0, k1=None):
    """C1: vacuum, 10-cell SC-PML on every side, z-polarised point dipole at the origin."""
    w = _common(N, LAMBDA / 20, (False, False, False), ((10,) * 3, (10,) * 3))
    k1 = N[2] if k1 is None else k1
    eps = np.zeros((N[0], N[1], k1 - k0, 3, 3), np.complex128)
    for v in range(3):
        eps[..., v, v] = 1.0
    w.update(eps=eps, k0=k0, k1=k1, full_eps=False, name=f"C1 vacuum box {N[0]}x{N[1]}x{N[2]} + 10-cell PML")
    return w


def c1_rhs(w):
    mdl = w["model"]
    fb.clear_srcs(mdl)
    fb.add_srce(mdl, fb.PointSrc([0.0, 0.0, 0.0], [0, 0, 1], 1.0))
    return fb.create_srcs(mdl)


def c2_waveguide(N=(200, 200, 200), k0=0, k1=None, period_z=None, seed=20261017, delta=20.0, npml=10):
    """C2: Si strip (eps 12.085, 500 x 220 nm, along x) in SiO2 (eps 2.085), 10-cell PML on all sides.
    Subpixel-smoothing stand-in: arithmetic/harmonic mixing on cells cut by the (axis-aligned) interfaces
    for the diagonal entries, plus a seeded random SYMMETRIC off-diagonal perturbation 0.2*(U-0.5) on
    the interface cells so the full-tensor path is exercised.  For weak scaling the cross-section repeats
    every `period_z` planes so that every z-slab carries the same work."""
    Nx, Ny, Nz = N
    w = _common(N, delta, (False, False, False), ((npml,) * 3, (npml,) * 3))
    k1 = Nz if k1 is None else k1
    period_z = Nz if period_z is None else period_z
    e_si, e_ox = 12.085, 2.085
    g = w["grid"]
    yp, yd = g.l[fb.PRIM][1], g.l[fb.DUAL][1]
    kk = np.arange(k0, k1)
    zc = ((kk % period_z) - period_z / 2.0) * delta          # primal z of each plane, periodic cross-section
    zp, zd = zc, zc + delta / 2
    eps = np.zeros((Nx, Ny, k1 - k0, 3, 3), np.complex128)
    hw_y, hw_z = 250.0, 110.0
    # E_x at (dual x, primal y, primal z); E_y at (primal, dual, primal); E_z at (primal, primal, dual)
    for v, (yy, zz) in enumerate(((yp, zp), (yd, zp), (yp, zd))):
        fy = _fill_1d(yy, delta, hw_y)[:, None]
        fz = _fill_1d(zz, delta, hw_z)[None, :]
        fill = fy * fz
        cut_y = (fy > 0) & (fy < 1) & (fz > 0)
        cut_z = (fz > 0) & (fz < 1) & (fy > 0)
        harmonic = cut_y if v == 1 else (cut_z if v == 2 else np.zeros_like(cut_y))
        eps[..., v, v] = _mix(fill, e_si, e_ox, harmonic)[None, :, :]
    # off-diagonal entries live at the voxel corners (primal,primal,primal)
    fy = _fill_1d(yp, delta, hw_y)[:, None]
    fz = _fill_1d(zp, delta, hw_z)[None, :]
    fill = fy * fz
    iface = ((fill > 0) & (fill < 1))[None, :, :]
    # include the cells next to the core faces so the perturbed region is a closed shell
    shell = np.zeros((Ny, k1 - k0), bool)
    inside = (fill >= 1)
    for s in (-1, 1):
        shell |= np.roll(inside, s, axis=0) & ~inside
        shell |= np.roll(inside, s, axis=1) & ~inside
    iface = iface | shell[None, :, :]
    rng = np.random.default_rng(seed + 1000 * k0)
    for (v, u) in ((0, 1), (0, 2), (1, 2)):
        pert = 0.2 * (rng.random((Nx, Ny, k1 - k0)) - 0.5) * iface
        eps[..., v, u] = pert
        eps[..., u, v] = pert
    w.update(eps=eps, k0=k0, k1=k1, full_eps=True,
             name=f"C2 Si strip waveguide in SiO2 {Nx}x{Ny}x{Nz}, full 3x3 eps, 10-cell PML")
    return w


def make_dense_offdiag(w, seed=7):
    """variant of a workload with a non-zero (symmetric) off-diagonal eps entry in EVERY cell: the fused full-tensor
    kernel path instead of diagonal kernel + correction pass"""
    rng = np.random.default_rng(seed)
    for (v, u) in ((0, 1), (0, 2), (1, 2)):
        pert = 0.05 * (rng.random(w["eps"].shape[:3]) - 0.5)
        w["eps"][..., v, u] = pert
        w["eps"][..., u, v] = pert
    w["full_eps"] = True
    w["name"] += " [dense off-diagonal variant]"
    return w


def c2_hh(N=(200, 200, 200)):
    """the C2 grid for the HH formulation A = Ce eps^-1 Cm - w^2 mu (model.jl:238-240): needs a diagonal eps, mu = 1"""
    w = c2_waveguide(N)
    for v in range(3):
        for u in range(3):
            if u != v:
                w["eps"][..., v, u] = 0
    w["full_eps"] = False
    w["name"] = f"C2 grid {N[0]}x{N[1]}x{N[2]}, HH formulation, diagonal eps, mu = 1"
    return w


def c2_rhs(w):
    """'mode-plane source' stand-in: y-polarised PlaneSrc at x = -1500 nm with a Gaussian transverse window."""
    mdl = w["model"]
    fb.clear_srcs(mdl)
    fb.add_srce(mdl, fb.PlaneSrc([1, 0, 0], -1500.0 if w["N"][0] >= 160 else 0.0, [0, 1, 0], 1.0))
    g = w["grid"]
    Y = g.l[fb.DUAL][1][None, :, None]
    Z = g.l[fb.PRIM][2][None, None, :]
    mdl.je_arr[..., 1] *= np.exp(-(Y / 300.0) ** 2 - (Z / 200.0) ** 2)
    return fb.create_srcs(mdl)


def c3_phc_slab(N=(256, 256, 128), k0=0, k1=None, seed=20261017, delta=15.0):
    """C3: photonic-crystal slab, Bloch-periodic in x,y (non-zero k), 10-cell PML in z; eps 12 slab of 16
    cells with an 8x8 supercell of air holes (a = Nx/8 cells, r = 0.3a); hole walls get the seeded
    symmetric off-diagonal perturbation (stand-in for Kottke smoothing on cylinders)."""
    Nx, Ny, Nz = N
    Lx, Ly = Nx * delta, Ny * delta
    kb = (0.30 * 2 * np.pi / Lx, 0.10 * 2 * np.pi / Ly, 0.0)
    w = _common(N, delta, (True, True, False), ((0, 0, 10), (0, 0, 10)), kb)
    k1 = Nz if k1 is None else k1
    g = w["grid"]
    a = Nx / 8.0 * delta
    r = 0.3 * a
    eps = np.zeros((Nx, Ny, k1 - k0, 3, 3), np.complex128)

    def hole_dist(x, y):
        X = (x[:, None] + Lx / 2) % a - a / 2
        Y = (y[None, :] + Ly / 2) % a - a / 2
        return np.sqrt(X * X + Y * Y)

    zpl = g.l[fb.PRIM][2][k0:k1]
    zdl = g.l[fb.DUAL][2][k0:k1]
    half_t = 8 * delta
    xp, xd = g.l[fb.PRIM][0], g.l[fb.DUAL][0]
    yp, yd = g.l[fb.PRIM][1], g.l[fb.DUAL][1]
    for v, (xx, yy, zz) in enumerate(((xd, yp, zpl), (xp, yd, zpl), (xp, yp, zdl))):
        d = hole_dist(xx, yy)
        fxy = np.clip((d - r) / delta + 0.5, 0.0, 1.0)          # 0 in hole, 1 in dielectric, linear ramp
        fz = _fill_1d(zz, delta, half_t)
        fill = fxy[:, :, None] * fz[None, None, :]
        eps[..., v, v] = 1.0 + 11.0 * fill
    d = hole_dist(xp, yp)
    wall = (np.abs(d - r) < delta)[:, :, None] & (_fill_1d(zpl, delta, half_t) > 0)[None, None, :]
    rng = np.random.default_rng(seed + 1000 * k0)
    for (v, u) in ((0, 1), (0, 2), (1, 2)):
        pert = 0.2 * (rng.random((Nx, Ny, k1 - k0)) - 0.5) * wall
        eps[..., v, u] = pert
        eps[..., u, v] = pert
    w.update(eps=eps, k0=k0, k1=k1, full_eps=True,
             name=f"C3 PhC slab {Nx}x{Ny}x{Nz}, Bloch x/y (k != 0), PML z, full 3x3 eps")
    return w


def c3_rhs(w):
    mdl = w["model"]
    fb.clear_srcs(mdl)
    zsrc = w["grid"].l[fb.PRIM][2][int(0.75 * w["N"][2])]
    fb.add_srce(mdl, fb.PlaneSrc([0, 0, 1], zsrc, [1, 0, 0], 1.0))
    return fb.create_srcs(mdl)


def c4_scatterer(N=(512, 512, 512), k0=0, k1=None, seed=20261017, delta=20.0, radius_cells=100):
    """C4: dielectric sphere (eps 4, radius 100 cells) in vacuum, 10-cell PML on all sides.  Smoothing stand-in:
    linear fill ramp across the surface for the diagonal entries + the seeded symmetric off-diagonal perturbation
    on the surface shell.  The eps array is built directly in Julia memory order ([u,v,k,j,i], C-contiguous) so
    that no transposing copy of the 19 GB array is needed."""
    Nx, Ny, Nz = N
    w = _common(N, delta, (False, False, False), ((10,) * 3, (10,) * 3))
    k1 = Nz if k1 is None else k1
    g = w["grid"]
    R = radius_cells * delta
    eps = np.zeros((3, 3, k1 - k0, Ny, Nx), np.complex128)
    xp, xd = g.l[fb.PRIM][0], g.l[fb.DUAL][0]
    yp, yd = g.l[fb.PRIM][1], g.l[fb.DUAL][1]
    zp, zd = g.l[fb.PRIM][2][k0:k1], g.l[fb.DUAL][2][k0:k1]

    def fill(x, y, z):
        d = np.sqrt(z[:, None, None] ** 2 + y[None, :, None] ** 2 + x[None, None, :] ** 2)
        return np.clip((R - d) / delta + 0.5, 0.0, 1.0)

    for v, (xx, yy, zz) in enumerate(((xd, yp, zp), (xp, yd, zp), (xp, yp, zd))):
        eps[v, v] = 1.0 + 3.0 * fill(xx, yy, zz)
    d = np.sqrt(zp[:, None, None] ** 2 + yp[None, :, None] ** 2 + xp[None, None, :] ** 2)
    shell = np.abs(d - R) < delta
    del d
    rng = np.random.default_rng(seed + 1000 * k0)
    for (v, u) in ((0, 1), (0, 2), (1, 2)):
        pert = np.zeros(shell.shape)
        pert[shell] = 0.2 * (rng.random(int(shell.sum())) - 0.5)
        eps[u, v] = pert
        eps[v, u] = pert
    w.update(eps=eps, k0=k0, k1=k1, full_eps=True, julia_layout=True,
             name=f"C4 dielectric sphere {Nx}x{Ny}x{Nz}, full 3x3 eps on the surface, 10-cell PML")
    return w


def c5_metalens(N=(1024, 1024, 96), k0=0, k1=None, seed=7, delta=20.0, pitch_cells=32):
    """C5 (weak-scaling unit: 96 z-planes per GPU, global Nz = 96 * n_gpus): metalens - eps 6.0 pillars with random
    radii (seed 7) on an eps 2.1 substrate, Bloch-periodic in x,y with k = 0, 10-cell PML in z.  Substrate = lower
    third of the global z range, pillars = the next 32 cells, air above.  Planar interfaces smooth to diagonal
    tensors; the pillar walls carry the seeded symmetric off-diagonal perturbation (Kottke stand-in).  Built in
    Julia memory order ([u,v,k,j,i]) so the 14.5 GB slab array needs no transposing copy."""
    Nx, Ny, Nz = N
    w = _common(N, delta, (True, True, False), ((0, 0, 10), (0, 0, 10)))
    k1 = Nz if k1 is None else k1
    g = w["grid"]
    a = pitch_cells * delta
    npx, npy = Nx // pitch_cells, Ny // pitch_cells
    radii = (6.0 + 7.0 * np.random.default_rng(seed).random((npx, npy))) * delta      # 6..13 cells
    Lx, Ly = Nx * delta, Ny * delta
    z0 = g.lg_prim[2][0]
    z_sub = z0 + (Nz // 3) * delta                    # substrate top
    z_top = z_sub + 32 * delta                        # pillar top

    def pillar(x, y):                                 # (signed distance to the wall, in-plane), shape (Ny, Nx)
        ix = np.minimum(((x + Lx / 2) // a).astype(int), npx - 1)
        iy = np.minimum(((y + Ly / 2) // a).astype(int), npy - 1)
        X = (x + Lx / 2) % a - a / 2
        Y = (y + Ly / 2) % a - a / 2
        return radii[ix[None, :], iy[:, None]] - np.sqrt(X[None, :] ** 2 + Y[:, None] ** 2)

    def slab_fill(z, lo, hi):                         # fraction of the cell [z-d/2, z+d/2] inside [lo, hi]
        return np.clip(np.minimum(z + delta / 2, hi) - np.maximum(z - delta / 2, lo), 0.0, delta) / delta

    eps = np.zeros((3, 3, k1 - k0, Ny, Nx), np.complex128)
    xp, xd = g.l[fb.PRIM][0], g.l[fb.DUAL][0]
    yp, yd = g.l[fb.PRIM][1], g.l[fb.DUAL][1]
    zp, zd = g.l[fb.PRIM][2][k0:k1], g.l[fb.DUAL][2][k0:k1]
    for v, (xx, yy, zz) in enumerate(((xd, yp, zp), (xp, yd, zp), (xp, yp, zd))):
        fxy = np.clip(pillar(xx, yy) / delta + 0.5, 0.0, 1.0)
        fsub = slab_fill(zz, -1e30, z_sub)
        fpil = slab_fill(zz, z_sub, z_top)
        eps[v, v] = 1.0 + 1.1 * fsub[:, None, None] + 5.0 * fpil[:, None, None] * fxy[None, :, :]
    wall_xy = np.abs(pillar(xp, yp)) < delta
    in_z = slab_fill(zp, z_sub, z_top) > 0
    rng = np.random.default_rng(seed + 1000 * k0 + 1)
    nwall = int(wall_xy.sum())
    for (v, u) in ((0, 1), (0, 2), (1, 2)):
        for kk in np.nonzero(in_z)[0]:
            pert = np.zeros((Ny, Nx))
            pert[wall_xy] = 0.2 * (rng.random(nwall) - 0.5)
            eps[u, v, kk] = pert
            eps[v, u, kk] = pert
    w.update(eps=eps, k0=k0, k1=k1, full_eps=True, julia_layout=True,
             name=f"C5 metalens {Nx}x{Ny}x{Nz} (96 planes per GPU), Bloch x/y, PML z, full 3x3 eps on pillar walls")
    return w


def tfsf_rhs(A, w, box_cells=140, axis=2, pol=0):
    """TF/SF plane-wave source stand-in for C4 (the reference has no TF/SF source, README.md:40), built from the
    operator itself:  b = A (M E_inc) - M (A E_inc),  with M the mask of the total-field region (a centred box of
    `box_cells` cells) and E_inc a plane wave travelling along +`axis`, polarised along `pol`, with the DISCRETE
    vacuum wavenumber (2/d) asin(w d / 2) so that A E_inc = 0 on the uniform part of the grid.  A is local, so b
    lives on the box surface only, and the solution of A e = b is E_inc + E_scat inside the box and E_scat
    outside.  `A` is anything with a matvec (`A @ x`) in the reference's DOF ordering."""
    assert axis != pol
    g = w["grid"]
    d = float(g.lg_prim[axis][1] - g.lg_prim[axis][0])
    k = 2.0 / d * np.arcsin(w["omega"] * d / 2.0)
    E = np.zeros(tuple(w["N"]) + (3,), np.complex128)
    inside = np.zeros(E.shape, bool)
    for v in range(3):
        pos = np.meshgrid(*[g.l[fb.DUAL if a == v else fb.PRIM][a] for a in range(3)], indexing="ij")
        if v == pol:
            E[..., v] = np.exp(-1j * k * pos[axis])
        half = box_cells / 2.0 * d
        inside[..., v] = (np.abs(pos[0]) < half) & (np.abs(pos[1]) < half) & (np.abs(pos[2]) < half)
    e_inc = fb.field_arr2vec(E)
    m = fb.field_arr2vec(inside.astype(np.complex128))
    return A @ (m * e_inc) - m * (A @ e_inc), m * e_inc


def make_operator(w, device=-1, rank=0, nranks=1, kernel=0, **kw):
    return fb.FdfdOperator(w["N"], w["isbloch"], w["sdl_e"], w["sdl_m"], w["omega"], w["eps"], None, w["e_mikL"],
                           device=device, rank=rank, nranks=nranks, kernel=kernel,
                           eps_has_offdiag=w["full_eps"], **kw)


# ---------------------------------------------------------------------------------------------------
# The same configurations described by OBJECTS (material pipeline): eps is rasterised and subpixel-
# smoothed on the device by the operator itself (fdfd_set_eps_objects) - no (Nx,Ny,Nz,3,3) host array.  These replace
# the analytic stand-ins above once the pipeline has been timed on hardware (round 2).
# ---------------------------------------------------------------------------------------------------
def c4_objects(N=(512, 512, 512), delta=20.0, radius_cells=100):
    """C4 from objects: vacuum box + one ball (eps 4, radius 100 cells) centred in the domain."""
    w = _common(N, delta, (False, False, False), ((10,) * 3, (10,) * 3))
    g = w["grid"]
    c = [0.5 * (g.bounds[0][a] + g.bounds[1][a]) for a in range(3)]
    shapes = [fb.Box(c, [2 * L for L in g.L]), fb.Ball(c, radius_cells * delta)]
    w.update(shapes=shapes, pinds=[0, 1], params=[np.eye(3), 4.0 * np.eye(3)], full_eps=True,
             name=f"C4 dielectric sphere {N[0]}x{N[1]}x{N[2]} from objects (Kottke-smoothed on the device), 10-cell PML")
    return w


def c5_objects(N=(1024, 1024, 96), seed=7, delta=20.0, pitch_cells=32):
    """C5 weak-scaling unit from objects: substrate slab (eps 2.1) + one cylinder (eps 6.0, seeded radius) per pitch cell."""
    w = _common(N, delta, (True, True, False), ((0, 0, 10), (0, 0, 10)))
    g = w["grid"]
    lo, L = g.bounds[0], g.L
    c = [lo[a] + 0.5 * L[a] for a in range(3)]
    zsub = lo[2] + 0.25 * L[2]
    shapes = [fb.Box(c, [2 * l for l in L]), fb.Box([c[0], c[1], lo[2] + 0.5 * (zsub - lo[2])], [2 * L[0], 2 * L[1], 0.5 * (zsub - lo[2])])]
    pinds = [0, 1]
    rng = np.random.default_rng(seed)
    pitch = pitch_cells * delta
    h = 0.2 * L[2]
    for j in range(N[1] // pitch_cells):
        for i in range(N[0] // pitch_cells):
            r = (0.15 + 0.25 * rng.random()) * pitch
            shapes.append(fb.Cylinder([lo[0] + (i + 0.5) * pitch, lo[1] + (j + 0.5) * pitch, zsub + h], r, h, axis=2))
            pinds.append(2)
    w.update(shapes=shapes, pinds=pinds, params=[np.eye(3), 2.1 * np.eye(3), 6.0 * np.eye(3)], full_eps=True,
             name=f"C5 metalens {N[0]}x{N[1]}x{N[2]} from objects ({len(shapes) - 2} pillars, Kottke-smoothed on the device)")
    return w


def make_operator_from_objects(w, device=-1, rank=0, nranks=1, kernel=0, **kw):
    A = fb.FdfdOperator(w["N"], w["isbloch"], w["sdl_e"], w["sdl_m"], w["omega"], None, None, w["e_mikL"],
                        device=device, rank=rank, nranks=nranks, kernel=kernel, **kw)
    A.set_eps_objects(w["grid"].lg_prim, w["shapes"], w["pinds"], w["params"])
    return A
