"""TEST INFRASTRUCTURE ONLY - multi-rank (z-slab) runs of the CPU logic-check build: one PROCESS per rank as on the GPU
box, the NCCL and driver entry points csrc/comm.cpp resolves with dlopen replaced by tests/emu/fakelibs (FIFOs between
the rank processes; stream memory operations as stores / bounded spins), "device" memory in named shared memory so
that the CUDA-IPC peer mapping of csrc/peer.cpp works between the processes.

    python tests/emu/run_emu_dist.py WORLD [apply] [krylov]        (parent: builds nothing, spawns the ranks)

Modes come from the environment the library itself reads (FDFD_PEER_DIRECT, FDFD_PEER_HALO, FDFD_INKERNEL_HALO_WAIT, FDFD_SPLIT_OVERLAP,
FDFD_NO_HALO_PREFETCH).  Every rank checks its own slab against the oracle on the global problem."""
import ctypes as C
import os
import shutil
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), HERE]


def parent(world, groups):
    emu = os.path.join(ROOT, "build", "emu")
    tmp = tempfile.mkdtemp(prefix="fdfd_emu_dist_")
    env = dict(os.environ, FDFD_B200_LIB=os.path.join(emu, "libfdfd_emu.so"), FDFD_EMU_IPC="1",
               LD_LIBRARY_PATH=os.path.join(emu, "fakelibs") + ":" + os.environ.get("LD_LIBRARY_PATH", ""),
               EMU_DIST_DIR=tmp, OMP_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--rank", str(r), str(world), *groups], env=env)
             for r in range(world)]
    deadline = time.time() + 600
    rc = 0
    for p in procs:
        try:
            rc |= p.wait(timeout=max(1, deadline - time.time()))
        except subprocess.TimeoutExpired:
            rc |= 1
            print("emu dist: a rank did not finish (dead-lock?)", flush=True)
            for q in procs:
                q.kill()
            break
    shutil.rmtree(tmp, ignore_errors=True)
    sys.exit(rc)


def wait_file(path, timeout=120):
    t0 = time.time()
    while not os.path.exists(path):
        if time.time() - t0 > timeout:
            raise TimeoutError(path)
        time.sleep(0.01)


def child(rank, world, groups):
    import numpy as np
    from problems import Problem, rel
    import maxwellfdm_jl_b200 as fb
    L = fb._lib
    assert "EMULATED" in L.lib().fdfd_version().decode()
    tmp = os.environ["EMU_DIST_DIR"]
    counter = [0]

    def comm_id():
        """rank 0 creates the unique id, the others read it from a file (the 128 bytes may travel by any means)"""
        counter[0] += 1
        path = os.path.join(tmp, f"uid_{counter[0]}")
        if rank == 0:
            uid = fb.comm_unique_id()
            with open(path + ".tmp", "wb") as f:
                f.write(uid)
            os.rename(path + ".tmp", path)
            return uid
        wait_file(path)
        return open(path, "rb").read()

    def slab_operator(p, **kw):
        k0, k1 = fb.partition(p.N[2], world, rank)
        A = fb.FdfdOperator(p.N, p.isbloch, p.sdl_e, p.sdl_m, p.omega, p.eps[:, :, k0:k1],
                            p.mu[:, :, k0:k1] if p.with_mu else None, p.ph, order_cmpfirst=p.cmpfirst,
                            boundft=["E" if b == 0 else "H" for b in p.boundft], ft="E" if p.ft == 0 else "H",
                            device=0, rank=rank, nranks=world, **kw)
        A.comm_init(comm_id())
        return A, k0, k1

    def slab_of(p, v, k0, k1):
        Nx, Ny, Nz = p.N
        if p.cmpfirst:
            return v[3 * Nx * Ny * k0:3 * Nx * Ny * k1].copy()
        return np.ascontiguousarray(v.reshape(3, Nz, Ny * Nx)[:, k0:k1]).ravel()

    def plane_of(p, v, k):
        """plane k of a global DOF vector in the halo layout (None outside a non-periodic grid)"""
        Nx, Ny, Nz = p.N
        if k < 0 or k >= Nz:
            if not p.isbloch[2]:
                return None
            k %= Nz
        if p.cmpfirst:
            return v[3 * Nx * Ny * k:3 * Nx * Ny * (k + 1)].copy()
        return np.ascontiguousarray(v.reshape(3, Nz, Ny * Nx)[:, k]).ravel()

    def host_halo_apply(A, p, x, k0, k1, transpose=False):
        """fdfd_apply_host_halos: the caller holds the whole vector, the neighbour planes ride along (no exchange)"""
        xs, lo, hi = slab_of(p, x, k0, k1), plane_of(p, x, k0 - 1), plane_of(p, x, k1)
        y = np.full(A.n, np.nan + 1j * np.nan)
        L.check(L.lib().fdfd_apply_host_halos(A._h, xs.ctypes.data, None if lo is None else lo.ctypes.data,
                                              None if hi is None else hi.ctypes.data, y.ctypes.data, 1 if transpose else 0), A._h)
        return y

    def dev_apply(A, x, transpose=False):
        y = np.full(A.n, np.nan + 1j * np.nan)
        f = L.lib().fdfd_apply_transpose if transpose else L.lib().fdfd_apply
        L.check(f(A._h, x.ctypes.data, y.ctypes.data, L.DEVICE), A._h)
        return y

    nchecks = 0
    if "apply" in groups:
        cases = []
        for isbloch in ((True, True, True), (False, True, False), (True, False, True)):
            for full, mu, cf, kern in ((True, True, True, 0), (False, False, True, 0), (True, False, False, 0), (True, True, True, 1)):
                cases.append(dict(N=(21, 18, 2 * world + 3), isbloch=isbloch, full_eps=full, with_mu=mu, cmpfirst=cf, kernel=kern))
        cases.append(dict(N=(9, 7, world), isbloch=(True, True, True), full_eps=True, with_mu=True, cmpfirst=True, kernel=0))
        for isbloch in ((True, True, True), (False, True, False)):
            cases.append(dict(N=(21, 18, 2 * world + 3), isbloch=isbloch, ft=1, full_mu=True, kernel=0))
            cases.append(dict(N=(33, 10, 3 * world + 1), isbloch=isbloch, boundft=(1, 1, 1), full_eps=True, with_mu=True, kernel=0))
        # slabs deep enough for >= 3 z-chunks (the in-kernel halo wait gates the first and last chunk) and several tiles
        cases.append(dict(N=(40, 14, 14 * world), isbloch=(True, False, True), full_eps=False, with_mu=False, kernel=0, lz=4))
        cases.append(dict(N=(40, 14, 14 * world), isbloch=(False, True, False), full_eps=True, with_mu=True, kernel=0, lz=3))
        # off-diagonal material on ONE slab only (every rank must still build and exchange the same arrays), once
        # pointwise symmetric on that slab and once not
        cases.append(dict(N=(21, 18, 4 * world), isbloch=(True, True, True), full_eps=True, with_mu=False, kernel=0, only_slab=1, sym=True))
        cases.append(dict(N=(21, 18, 4 * world), isbloch=(False, True, False), full_eps=True, with_mu=True, kernel=0, only_slab=0, sym=False))
        # slabs of >= 16 planes: the host-buffer applies run the sub-slab pipeline (per slab, halos from the host vector)
        cases.append(dict(N=(21, 10, 17 * world), isbloch=(True, False, True), full_eps=True, with_mu=False, kernel=0))
        cases.append(dict(N=(21, 10, 16 * world + 1), isbloch=(False, True, False), full_eps=False, with_mu=False, kernel=0, real_mass=True))
        for cs in cases:
            kern = cs.pop("kernel")
            lz = cs.pop("lz", 0)
            only_slab, sym = cs.pop("only_slab", None), cs.pop("sym", False)
            if lz:
                os.environ["FDFD_LZ"] = str(lz)
            else:
                os.environ.pop("FDFD_LZ", None)
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
            mf = p.oracle_matfree()
            x = p.random_x()
            A, k0, k1 = slab_operator(p, kernel=kern)
            if only_slab is not None:
                assert A.offdiag_symmetric == sym, (rank, sym)
            xs = slab_of(p, x, k0, k1)
            for rep in range(3):             # back-to-back applies: epochs of the halo protocol advance
                y = dev_apply(A, xs)
            e1 = rel(y, slab_of(p, mf(x), k0, k1))
            yh = A @ xs                      # host-buffer path
            e2 = rel(yh, slab_of(p, mf(x), k0, k1))
            yt = dev_apply(A, xs, True)
            A_ref, _ = p.oracle_csc()
            yT_ref = slab_of(p, A_ref.to_scipy().T @ x, k0, k1)
            e3 = rel(yt, yT_ref)
            e4 = rel(host_halo_apply(A, p, x, k0, k1), slab_of(p, mf(x), k0, k1))
            e5 = rel(host_halo_apply(A, p, x, k0, k1, True), yT_ref)
            e6 = rel(dev_apply(A, xs), slab_of(p, mf(x), k0, k1))      # and the exchange path still works afterwards
            A.close()
            assert max(e1, e2, e3, e4, e5, e6) < 1e-12, (rank, cs, kern, e1, e2, e3, e4, e5, e6)
            nchecks += 6
    if "apply" in groups:
        # eps from objects on z-slabs: every rank rasterises its own planes; the slabs together are the single-slab operator
        from problems import matparams_scene
        from oracle.grid import EE
        os.environ.pop("FDFD_LZ", None)
        N, isbloch = (14, 11, 3 * world + 2), (True, False, True)
        lp, _, f_sh, pinds, params = matparams_scene(N, isbloch, False, 6, False)
        p = Problem(N, isbloch)
        eps = np.zeros(N + (3, 3), complex)
        k0, k1 = fb.partition(N[2], world, rank)
        A = fb.FdfdOperator(N, isbloch, p.sdl_e, p.sdl_m, p.omega, None, None, p.ph, device=0, rank=rank, nranks=world)
        A.set_eps_objects(lp, f_sh, pinds, params)
        A.comm_init(comm_id())
        # reference: the array of the whole grid (same kernel, so the comparison is exact up to the operator's own rounding)
        p.eps = fb.calc_matparams_array(fb.Grid(lp, isbloch), (EE,) * 3, EE, f_sh, pinds, params, device=0)
        p.full_eps = True
        x = p.random_x()
        e = rel(dev_apply(A, slab_of(p, x, k0, k1)), slab_of(p, p.oracle_matfree()(x), k0, k1))
        assert e < 1e-12 and A.offdiag_symmetric, (rank, e)
        A.close()
        nchecks += 1
    if "krylov" in groups:
        os.environ.pop("FDFD_LZ", None)
        import scipy.sparse.linalg as spla
        for ft, kw in ((0, dict(full_eps=True)), (1, dict(with_mu=True))):
            p = Problem((12, 10, 4 * world + 1), (True, False, True), ft=ft, omega=1.3 - 0.4j, **kw)
            A_ref, _ = p.oracle_csc()
            b = A_ref.matvec(p.random_x(5))
            x_ref = spla.splu(A_ref.to_scipy().tocsc()).solve(b)
            A, k0, k1 = slab_operator(p, kernel=2)
            bs = slab_of(p, b, k0, k1)
            for method in (L.BICGSTAB, L.QMR):
                x = np.zeros(A.n, complex)
                iters, relres = C.c_int(), C.c_double()
                code = L.lib().fdfd_solve(A._h, method, bs.ctypes.data, x.ctypes.data, L.DEVICE, 1e-10, 400, 10,
                                          C.byref(iters), C.byref(relres), None)
                L.check(code, A._h)
                e = rel(x, slab_of(p, x_ref, k0, k1))
                assert e < 1e-7, (rank, ft, method, e, iters.value, relres.value)
                nchecks += 1
            # five BiCGSTAB iterations against the textbook iteration on the global operator: inner products that miss a
            # slab's share (sigma is accumulated by the three launches that write p) keep r = b - A x consistent and
            # may still converge, so only the trajectory shows them
            from fuzz_krylov_emu import bicgstab_ref
            x = np.zeros(A.n, complex)
            code = L.lib().fdfd_solve(A._h, L.BICGSTAB, bs.ctypes.data, x.ctypes.data, L.DEVICE, 1e-300, 5, 1,
                                      C.byref(iters), C.byref(relres), None)
            assert code in (L.OK, L.ENOCONV), (rank, code)
            e = rel(x, slab_of(p, bicgstab_ref(p.oracle_matfree(), b, 5), k0, k1))
            assert e < 1e-9, (rank, ft, "trajectory", e)
            nchecks += 1
            A.close()
    modes = [k for k in ("FDFD_PEER_DIRECT", "FDFD_PEER_HALO", "FDFD_INKERNEL_HALO_WAIT", "FDFD_SPLIT_OVERLAP", "FDFD_NO_HALO_PREFETCH") if os.environ.get(k)]
    print(f"emu dist rank {rank}/{world} [{','.join(modes) or 'default'}]: {nchecks} checks ok", flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "--rank":
        child(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4:] or ["apply", "krylov"])
    else:
        parent(int(sys.argv[1]), sys.argv[2:] or ["apply", "krylov"])
