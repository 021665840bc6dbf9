"""2+ GPU check of the experimental z-slab paths: slab results must equal the single-slab GPU operator, first with a
host sync after every apply, then (--stress) as a back-to-back stream of applies.
    default        FDFD_INKERNEL_HALO_WAIT=1  (single launch, boundary z-chunks gated on a flag)
    --peer         + FDFD_PEER_HALO=1         (SM-free exchange over IPC-mapped peer memory, csrc/peer.cpp)
    --peer-only    FDFD_PEER_HALO=1 alone     (exchange, then launch)
Run under `timeout`: round 1 saw the default mode stall in the back-to-back stream."""
import os
import sys

if "--peer-only" not in sys.argv:
    os.environ["FDFD_INKERNEL_HALO_WAIT"] = "1"
if "--peer" in sys.argv or "--peer-only" in sys.argv:
    os.environ["FDFD_PEER_HALO"] = "1"
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
    for isbloch, full in (((True, False, True), True), ((False, True, False), False), ((True, True, True), True)):
        p = Problem((70, 45, 40 * world), isbloch, full_eps=full)
        if full:                         # sparse off-diagonals: diagonal kernel + correction pass
            p.eps[:, :, 30:, 0, 1] = p.eps[:, :, 30:, 1, 0] = 0
        x = p.random_x()
        k0, k1 = fb.partition(p.N[2], world, rank)
        A = fb.FdfdOperator(p.N, p.isbloch, p.sdl_e, p.sdl_m, p.omega, p.eps[:, :, k0:k1], None, p.ph,
                            device=local, rank=rank, nranks=world)
        uid = [fb.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        A.comm_init(uid[0])
        n3 = 3 * p.N[0] * p.N[1]
        xs = torch.from_numpy(x[n3 * k0:n3 * k1].copy()).cuda()
        ys = [A @ xs for _ in range(3)]                       # three epochs of the flag
        yt = A.rmatvec_T(xs)
        parts = [None] * world
        dist.all_gather_object(parts, (k0, ys[2].cpu().numpy(), yt.cpu().numpy()))
        if not (torch.equal(ys[0], ys[1]) and torch.equal(ys[1], ys[2])):
            fails.append(("not reproducible", isbloch))
        if rank == 0:
            y = np.concatenate([a[1] for a in sorted(parts, key=lambda t: t[0])])
            ytg = np.concatenate([a[2] for a in sorted(parts, key=lambda t: t[0])])
            A1 = p.operator(device=local)
            e1, e2 = rel(y, A1 @ x), rel(ytg, A1.rmatvec_T(x))
            if not (e1 < 1e-13 and e2 < 1e-13):
                fails.append((isbloch, e1, e2))
            A1.close()
        if "--stress" in sys.argv:
            if rank == 0:
                print("stress: 50 back-to-back applies ...", flush=True)
            ys2 = torch.empty_like(xs)
            tot, _ = A.bench_apply(xs, ys2, warmup=3, iters=50)
            if not torch.equal(ys2, ys[2]):
                fails.append(("back-to-back result differs", isbloch))
            if rank == 0:
                print(f"stress ok: {tot / 50 * 1e3:.1f} us per apply", flush=True)
        A.close()
    flag = torch.tensor([len(fails)], device="cuda")
    dist.all_reduce(flag)
    if rank == 0:
        print("INKERNEL_HALO_CHECK", "FAIL" if flag.item() else "OK", "world", world, fails[:4])
    dist.destroy_process_group()
    sys.exit(1 if flag.item() else 0)


if __name__ == "__main__":
    main()
