"""World-size-2 (and 3) checks of the z-slab decomposition on CPU with the gloo backend.

Covers the host-side logic of the N>1 path: the partition rule (fdfd_partition), the halo plan and message
order the library uses with NCCL (fdfd_halo_plan; the P == 2 Bloch case where both neighbours are the
same rank), and the claim the multi-GPU design rests on: with ONE halo plane on each side
a rank can compute its own rows of y = A x.  The operator rows come from the oracle's CSC.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from problems import Problem, rel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, wrapz, N, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import maxwellfdm_jl_b200 as fb
        p = Problem(N, (True, False, wrapz), full_eps=True, with_mu=True)
        A, _ = p.oracle_csc()
        S = A.to_scipy().tocsr()
        x = p.random_x()
        y_ref = S @ x
        Nx, Ny, Nz = N
        plane = 3 * Nx * Ny                                  # cmp-first: a z-plane is a contiguous block
        k0, k1 = fb.partition(Nz, world, rank)
        up, dn = fb.halo_plan(world, rank, wrapz)
        own = torch.from_numpy(x[k0 * plane:k1 * plane].copy())
        lo = torch.zeros(plane, dtype=torch.complex128)
        hi = torch.zeros(plane, dtype=torch.complex128)
        first, last = own[:plane].clone(), own[-plane:].clone()
        # same message order as comm.cpp: {last -> up, lo <- dn} then {first -> dn, hi <- up}
        ops = []
        if up >= 0:
            ops.append(dist.P2POp(dist.isend, torch.view_as_real(last), up))
        if dn >= 0:
            ops.append(dist.P2POp(dist.irecv, torch.view_as_real(lo), dn))
        if dn >= 0:
            ops.append(dist.P2POp(dist.isend, torch.view_as_real(first), dn))
        if up >= 0:
            ops.append(dist.P2POp(dist.irecv, torch.view_as_real(hi), up))
        for r in (dist.batch_isend_irecv(ops) if ops else []):
            r.wait()
        # the halos must be the wrapped neighbour planes (or untouched zeros at a symmetry end)
        klo, khi = (k0 - 1) % Nz, k1 % Nz
        exp_lo = x[klo * plane:(klo + 1) * plane] if dn >= 0 else np.zeros(plane)
        exp_hi = x[khi * plane:(khi + 1) * plane] if up >= 0 else np.zeros(plane)
        ok_halo = np.array_equal(lo.numpy(), exp_lo) and np.array_equal(hi.numpy(), exp_hi)
        # own rows of A only touch planes k0-1 .. k1: rebuild x from own + halos, zero elsewhere
        x_loc = np.zeros_like(x)
        x_loc[k0 * plane:k1 * plane] = own.numpy()
        x_loc[klo * plane:(klo + 1) * plane] = lo.numpy() if dn >= 0 else x_loc[klo * plane:(klo + 1) * plane]
        x_loc[khi * plane:(khi + 1) * plane] = hi.numpy() if up >= 0 else x_loc[khi * plane:(khi + 1) * plane]
        if dn < 0 and world > 1:
            pass  # symmetry end: the plane below does not exist; the coefficients there are zero
        rows = slice(k0 * plane, k1 * plane)
        y_loc = S[rows] @ x_loc
        err = rel(y_loc, y_ref[rows])
        # allreduce of a Krylov-style inner product (sum over slabs == global dot)
        d = torch.tensor([np.vdot(own.numpy(), own.numpy()).real])
        dist.all_reduce(d)
        ok_dot = abs(d.item() - np.vdot(x, x).real) < 1e-9 * d.item()
        q.put((rank, ok_halo, err, ok_dot, (k0, k1, up, dn)))
        dist.destroy_process_group()
    except Exception as e:  # noqa: BLE001
        q.put((rank, False, repr(e), False, None))


@pytest.mark.parametrize("world,wrapz,N", [(2, True, (5, 4, 7)), (2, False, (5, 4, 7)), (3, True, (4, 3, 8)),
                                           (2, True, (3, 3, 2))])
def test_slab_halo_plan_over_gloo(world, wrapz, N):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, wrapz, N, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
    for rank, ok_halo, err, ok_dot, info in sorted(res):
        assert ok_halo, (rank, info, err)
        assert isinstance(err, float) and err < 1e-13, (rank, info, err)
        assert ok_dot


def test_halo_plan_rule():
    import maxwellfdm_jl_b200 as fb
    assert fb.halo_plan(1, 0, True) == (0, 0) and fb.halo_plan(1, 0, False) == (-1, -1)
    assert fb.halo_plan(2, 0, True) == (1, 1) and fb.halo_plan(2, 1, True) == (0, 0)
    assert fb.halo_plan(4, 0, False) == (1, -1) and fb.halo_plan(4, 3, False) == (-1, 2)
    assert fb.halo_plan(4, 3, True) == (0, 2)
    with pytest.raises(ValueError):
        fb.halo_plan(2, 2, True)
