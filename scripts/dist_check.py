"""Multi-GPU (z-slab) correctness check, launched with torch.distributed.run, one rank per GPU.

Every rank builds the same seeded global problem, creates its slab operator (NCCL halo exchange +
allreduce inside libfdfd_b200.so), applies / solves on its slab, and rank 0 compares the gathered result
with the CPU oracle (small grids) and with a single-slab GPU operator (larger grid)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist

from problems import Problem, rel
import maxwellfdm_jl_b200 as fb


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    fails = []

    def slab_operator(p, **kw):
        k0, k1 = fb.partition(p.N[2], world, rank)
        A = fb.FdfdOperator(p.N, p.isbloch, p.sdl_e, p.sdl_m, p.omega, p.eps[:, :, k0:k1],
                            p.mu[:, :, k0:k1] if p.with_mu else None, p.ph, order_cmpfirst=p.cmpfirst,
                            boundft=["E" if b == 0 else "H" for b in p.boundft], ft="E" if p.ft == 0 else "H",
                            device=local, rank=rank, nranks=world, **kw)
        uid = [fb.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        A.comm_init(uid[0])
        return A, k0, k1

    def slab_of(p, v, k0, k1):
        """this rank's slab of a global DOF vector"""
        Nx, Ny, Nz = p.N
        if p.cmpfirst:
            return v[3 * Nx * Ny * k0:3 * Nx * Ny * k1].copy()
        return np.ascontiguousarray(v.reshape(3, Nz, Ny * Nx)[:, k0:k1]).ravel()

    def gather(p, ys, k0, k1):
        parts = [None] * world
        dist.all_gather_object(parts, (k0, k1, ys))
        Nx, Ny, Nz = p.N
        if p.cmpfirst:
            return np.concatenate([a[2] for a in sorted(parts, key=lambda t: t[0])])
        out = np.empty((3, Nz, Ny * Nx), complex)
        for a0, a1, a in parts:
            out[:, a0:a1] = a.reshape(3, a1 - a0, Ny * Nx)
        return out.ravel()

    cases = []
    for isbloch in ((True, True, True), (False, True, False), (True, False, True)):
        for full, mu, cf, kern in ((True, True, True, 0), (False, False, True, 0), (True, False, False, 0),
                                   (True, True, True, 1)):
            cases.append(dict(N=(21, 18, 2 * world + 3), isbloch=isbloch, full_eps=full, with_mu=mu, cmpfirst=cf,
                              kernel=kern))
    cases.append(dict(N=(9, 7, world), isbloch=(True, True, True), full_eps=True, with_mu=True, cmpfirst=True, kernel=0))
    # mirrored arrangement across slabs: HH formulation (full-tensor mu), EE on boundft all-HH (full eps)
    for isbloch in ((True, True, True), (False, True, False)):
        cases.append(dict(N=(21, 18, 2 * world + 3), isbloch=isbloch, ft=1, full_mu=True, kernel=0))
        cases.append(dict(N=(33, 10, 3 * world + 1), isbloch=isbloch, boundft=(1, 1, 1), full_eps=True, with_mu=True,
                          kernel=0))
    # off-diagonal material on ONE slab only (every rank must still build and exchange the same arrays), once
    # pointwise symmetric on that slab and once not
    cases.append(dict(N=(21, 18, 4 * world), isbloch=(True, True, True), full_eps=True, with_mu=False, kernel=0, only_slab=1, sym=True))
    cases.append(dict(N=(21, 18, 4 * world), isbloch=(False, True, False), full_eps=True, with_mu=True, kernel=0, only_slab=0, sym=False))
    for cs in cases:
        kern = cs.pop("kernel")
        only_slab, sym = cs.pop("only_slab", None), cs.pop("sym", False)
        p = Problem(**cs)
        if only_slab is not None:
            import itertools
            a0, a1 = fb.partition(p.N[2], world, only_slab % world)
            for v, u in itertools.permutations(range(3), 2):
                p.eps[:, :, :a0, v, u] = 0
                p.eps[:, :, a1:, v, u] = 0
            if sym:
                for v, u in itertools.combinations(range(3), 2):
                    p.eps[..., u, v] = p.eps[..., v, u]
        A_ref, _ = p.oracle_csc()
        x = p.random_x()
        A, k0, k1 = slab_operator(p, kernel=kern)
        if only_slab is not None and A.offdiag_symmetric != sym:
            fails.append(("offdiag_symmetric", cs, sym))
        y = gather(p, A @ slab_of(p, x, k0, k1), k0, k1)
        e1 = rel(y, A_ref.matvec(x))
        yt = gather(p, A.rmatvec_T(slab_of(p, x, k0, k1)), k0, k1)
        e2 = rel(yt, A_ref.to_scipy().T @ x)
        if not (e1 < 1e-12 and e2 < 1e-12):
            fails.append(("apply", cs, kern, e1, e2))
        A.close()

    # deep slabs (>= 16 planes per rank, several z-chunks per tile column): the host-buffer apply runs the sub-slab pipeline
    # (boundary planes up first, exchanged on the device), the device apply overlaps the exchange with the interior z-chunks
    # (in-kernel halo wait) when the peer exchange is the data plane; back-to-back applies advance the halo epochs
    for isbloch, kw in (((True, False, True), dict(full_eps=True)),
                        ((False, True, False), dict(full_eps=True, real_mass=True, sym_real_off=True)),
                        ((True, True, True), dict(full_eps=False, real_mass=True))):
        p = Problem((40, 31, 18 * world + 1), isbloch, **kw)
        x = p.random_x()
        y_ref = p.oracle_matfree()(x)
        A, k0, k1 = slab_operator(p)
        xs = slab_of(p, x, k0, k1)
        xd = torch.from_numpy(xs).cuda()
        for rep in range(3):
            yd = A @ xd
        e1 = rel(gather(p, yd.cpu().numpy(), k0, k1), y_ref)
        for rep in range(2):
            yh = A @ xs
        e2 = rel(gather(p, yh, k0, k1), y_ref)
        e3 = rel(gather(p, (A @ xd).cpu().numpy(), k0, k1), y_ref)     # device path again after host applies
        if not (e1 < 1e-12 and e2 < 1e-12 and e3 < 1e-12):
            fails.append(("deep slabs", isbloch, kw, A.halo_data_plane, e1, e2, e3))
        A.close()
    cases = cases + ["deep"] * 3

    # Krylov across slabs (allreduced dots): vacuum-like PML box with a point-like right-hand side
    from oracle.grid import Grid, create_stretched_dls
    n, nz = 16, max(16, 4 * world)
    lam = 8.0
    omega = 2 * np.pi / lam
    grid = Grid(((np.arange(n + 1) - n / 2) * 1.0, (np.arange(n + 1) - n / 2) * 1.0, (np.arange(nz + 1) - nz / 2) * 1.0),
                (False, False, False))
    sdl_e, sdl_m, sei, smi = create_stretched_dls(omega, grid, ((4,) * 3, (4,) * 3))
    eps = np.zeros(grid.N + (3, 3), complex)
    for v in range(3):
        eps[..., v, v] = 1.0
    from oracle import operators as oop
    Ce, Cm = oop.create_curls(sei, smi, (0, 0, 0), grid.isbloch, np.ones(3, complex))
    Pe, Pm = oop.create_paramops(eps, np.broadcast_to(np.eye(3), grid.N + (3, 3)), sdl_e, sdl_m, sei, smi, (0, 0, 0),
                                 grid.isbloch, np.ones(3, complex))
    A_ref = oop.create_A(0, omega, Pe, Pm, Ce, Cm)
    b = np.zeros(A_ref.shape[0], complex)
    b[3 * (n // 2 + n * (n // 2 + n * (nz // 2))) + 2] = 1.0

    class PB:
        N, cmpfirst = grid.N, True
    k0, k1 = fb.partition(nz, world, rank)
    A = fb.FdfdOperator(grid.N, grid.isbloch, sdl_e, sdl_m, omega, eps[:, :, k0:k1], None, np.ones(3, complex),
                        device=local, rank=rank, nranks=world)
    uid = [fb.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    A.comm_init(uid[0])
    for method in ("bicgstab", "qmr"):
        xs, info = A.solve(torch.from_numpy(slab_of(PB, b, k0, k1)).cuda(), method=method, rtol=1e-9, maxit=6000,
                           check_every=10)
        xg = gather(PB, xs.cpu().numpy(), k0, k1)
        res = rel(A_ref.matvec(xg), b)
        if not (info["converged"] and res < 1e-8):
            fails.append(("solve", method, info, res))
    A.close()

    # larger grid: slab result == single-slab GPU result (tiled kernel, several tiles and chunks)
    p = Problem((70, 45, 8 * world + 5), (True, False, True), full_eps=True)
    x = p.random_x()
    A, k0, k1 = slab_operator(p)
    y = gather(p, A @ slab_of(p, x, k0, k1), k0, k1)
    A.close()
    if rank == 0:
        A1 = p.operator(device=local)
        e = rel(y, A1 @ x)
        if not e < 1e-13:
            fails.append(("vs single slab", e))
        A1.close()
    flag = torch.tensor([len(fails)], device="cuda")
    dist.all_reduce(flag)
    if rank == 0:
        print("DIST_CHECK", "FAIL" if flag.item() else "OK", "world", world, "cases", len(cases), fails[:5])
    dist.destroy_process_group()
    sys.exit(1 if flag.item() else 0)


if __name__ == "__main__":
    main()
