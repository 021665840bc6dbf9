#!/usr/bin/env python
"""bench.py - FDFD operator throughput (GDOF/s) and Krylov iterations/s on B200, BASELINE.json metric.

A "step" is ONE application y = A x of the matrix-free operator A = curl mu^-1 curl - w^2 eps on the workload of
BASELINE.json configs[1] (C2: 200^3 Si strip waveguide, full 3x3 eps, 10-cell PML); with N GPUs the grid is
200 x 200 x (200 N) split into N z-slabs (weak scaling, one NCCL halo exchange per apply).
    value        GDOF/s, x / eps resident in HBM, CUDA events on the library's stream, max over ranks
    e2e          the same metric through the C-ABI call with HOST (pinned) buffers: H2D of x and D2H of y inside the
                 timed region;  e2e_solve: fdfd_solve(FDFD_HOST), b in / x out, fixed 200 BiCGSTAB iterations
    roofline     algorithmic bytes per DOF (x 16 + y 16 + eps_diag 16, or 8 when the diagonal mass entries are real and
                 travel as doubles; + the off-diagonal streams on the blocks that hold any) / apply time against the
                 measured HBM copy bandwidth (MEASURED_PEAKS.json)
    parity       UNTIMED check before the timed region, at every N: the N-slab operator and 5 BiCGSTAB iterations on
                 a reduced copy of the workload against the CPU oracle (oracle/ is the checker, never the thing timed)
    configs      (N = 1) the other BASELINE configurations: C1, C3, C4 512^3, C5 unit slab, dense off-diagonals, HH
    scale_c5 / scale_c4   C5 weak scaling (1024 x 1024 x 96 per GPU) and C4 strong scaling (512^3 / N) at this N
    halo         (N > 1) the halo exchange by itself: bytes, microseconds, fraction of the NVLink peer bandwidth
    cpu_baseline the oracle's Julia-style single-thread CSC mul! on a bounded sample of the workload
--impl reference times the reference's CPU path stand-in (CSC assembled by the oracle restatement; product and an
unpreconditioned BiCGSTAB on all host cores) - the reference itself is Julia and cannot run here.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fdfd_operator_apply_throughput"
UNIT = "GDOF/s"
PER_GPU_N = (200, 200, 200)
SAMPLE_PLANES = 8
NVLINK_PEER_GBS = 770.0        # measured peer copy per direction on this pool (B200_PROFILING.md); nominal 900


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled through NVML every ~2 ms while the timed region runs (an nvidia-smi
    subprocess is too slow for a region of a few ms).  start_and_wait() returns once the first sample exists, and a
    sample is taken synchronously at stop, so even a 5 ms region is bracketed."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.err = index, [], False, None
        self.max_mhz = None
        self.ready = threading.Event()
        self._h = self._nv = self._reasons = None

    def _open(self):
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        try:       # honour CUDA_VISIBLE_DEVICES-style remapping by matching the torch device's UUID when possible
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            for i in range(nv.nvmlDeviceGetCount()):
                hh = nv.nvmlDeviceGetHandleByIndex(i)
                u = nv.nvmlDeviceGetUUID(hh)
                u = u.decode() if isinstance(u, bytes) else u
                if uuid in u:
                    h = hh
                    break
        except Exception:
            pass
        self._nv, self._h = nv, h
        self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        self._reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            nv.nvmlDeviceGetCurrentClocksThrottleReasons

    def _sample(self):
        nv, h = self._nv, self._h
        self.samples.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(self._reasons(h))))

    def run(self):
        try:
            self._open()
            self._sample()
            self.ready.set()
            while not self.stop_flag:
                self._sample()
                time.sleep(0.002)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
            self.ready.set()

    def start_and_wait(self):
        self.start()
        self.ready.wait(timeout=10)
        return self

    def summary(self):
        self.stop_flag = True
        self.join(timeout=2)
        try:
            if self._h is not None:
                self._sample()
        except Exception:
            pass
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable: " + str(self.err)]}
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                "hw_power_brake_slowdown": 0x80}
        allr = 0
        for _, r in self.samples:
            allr |= r
        return {"sm_mhz": statistics.median(c for c, _ in self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": [n for n, b in bits.items() if allr & b], "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------
# CPU legs (oracle = checker / reported baseline; never part of the product path)
# ---------------------------------------------------------------------------------------------------
def sample_workload():
    """Bounded sample of the C2 workload for the CPU legs: the SAMPLE_PLANES z-planes through the core."""
    import workloads
    Nx, Ny, Nz = PER_GPU_N
    k0 = Nz // 2 - SAMPLE_PLANES // 2
    w = workloads.c2_waveguide(PER_GPU_N, k0, k0 + SAMPLE_PLANES)
    return w, k0


def oracle_sample_matrix():
    """CSC of the sample slab assembled by the ORACLE (treated as a periodic-in-z stack of the 8 planes)."""
    import numpy as np
    from oracle import operators as op
    w, k0 = sample_workload()
    sl = slice(k0, k0 + SAMPLE_PLANES)
    sdl_e = (w["sdl_e"][0], w["sdl_e"][1], w["sdl_e"][2][sl])
    sdl_m = (w["sdl_m"][0], w["sdl_m"][1], w["sdl_m"][2][sl])
    sei = tuple(1 / a for a in sdl_e)
    smi = tuple(1 / a for a in sdl_m)
    isbloch = (False, False, True)
    ph = np.ones(3, complex)
    mu = np.zeros(w["eps"].shape, complex)
    for v in range(3):
        mu[..., v, v] = 1
    Ce, Cm = op.create_curls(sei, smi, (0, 0, 0), isbloch, ph)
    Pe, Pm = op.create_paramops(w["eps"], mu, sdl_e, sdl_m, sei, smi, (0, 0, 0), isbloch, ph)
    A = op.create_A(0, w["omega"], Pe, Pm, Ce, Cm)
    return A, f"{PER_GPU_N[0]}x{PER_GPU_N[1]}x{SAMPLE_PLANES} planes through the core of the C2 workload " \
              f"({A.shape[0]} DOF, nnz {A.nnz}), periodic in z"


def cpu_baseline_port(reps=5):
    import numpy as np
    from oracle import cbaseline as cb
    A, desc = oracle_sample_matrix()
    x = np.random.default_rng(1).standard_normal(A.shape[0]) + 0j
    y = np.empty_like(x)
    cb.csc_mul_serial(A, x, y)
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        cb.csc_mul_serial(A, x, y)
        ts.append(time.perf_counter() - t)
    return {"value": A.shape[0] / min(ts) / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": desc + f"; Julia-style single-thread CSC mul!, best of {reps}"}


def run_reference(args):
    """--impl reference: the reference's CPU path (stand-in) on all host cores; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import numpy as np
    from oracle import cbaseline as cb
    cb.use_all_cores()                       # torchrun pins OMP_NUM_THREADS=1; the reference arm uses every core
    A, desc = oracle_sample_matrix()
    R = cb.CsrOmp(A)
    n = A.shape[0]
    x = np.random.default_rng(1).standard_normal(n) + 1j * np.random.default_rng(2).standard_normal(n)
    y = np.empty_like(x)
    for _ in range(max(args.warmup, 1)):
        R.mul(x, y)
    ts = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        R.mul(x, y)
        ts.append(time.perf_counter() - t0)
    dt = sum(ts) / len(ts)
    cores = cb.num_threads()
    val = n / dt / 1e9
    # the solve half of the metric: unpreconditioned BiCGSTAB on the same CSC matrix, all
    # host cores, fixed iteration count (no convergence exit), as the GPU arm's krylov block
    kit = max(10, min(args.krylov_iters, 50))
    b = R.mul(x, np.empty_like(x)).copy()
    xs = np.zeros(n, np.complex128)
    R.bicgstab(b, xs, 3)                      # warm-up
    xs[:] = 0
    t0 = time.perf_counter()
    relres = R.bicgstab(b, xs, kit)
    tk = time.perf_counter() - t0
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "ms_per_step_min": min(ts) * 1e3,
            "ms_per_step_median": statistics.median(ts) * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "c128 (complex fp64)", "data": "synthetic",
            "config": {"workload": "C2 Si strip waveguide 200x200x200, full 3x3 eps, 10-cell PML",
                       "sample": desc, "l2": "matrix stream larger than LLC"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": desc + "; OpenMP CSR product of the oracle-assembled matrix (stand-in for the "
                                              "Julia SparseMatrixCSC mul!, which cannot run here: no julia binary)"},
            "krylov": {"method": "bicgstab", "iters": kit, "iter_per_s": kit / tk, "relres_after": relres,
                       "dof": n, "gdof_iter_per_s": n * kit / tk / 1e9,
                       "note": "unpreconditioned BiCGSTAB on the sampled CSC matrix, OpenMP on all host cores; compare "
                               "per DOF: iterations/s scale inversely with the DOF count"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# GPU legs
# ---------------------------------------------------------------------------------------------------
class Ctx:
    pass


def rand_vec(n, gen):
    import torch
    return torch.randn(n, 2, device="cuda", dtype=torch.float64, generator=gen).view(torch.complex128).reshape(-1)


def max_over_ranks(v, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(v)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(world):
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def build_operator(w, local, rank, world, **kw):
    import torch.distributed as dist
    import workloads
    import maxwellfdm_jl_b200 as fb
    if "shapes" in w:
        A = None
        if world > 1:
            # objects are rasterised on the first use, which needs the communicator (ghost planes): create, connect, set
            A = fb.FdfdOperator(w["N"], w["isbloch"], w["sdl_e"], w["sdl_m"], w["omega"], None, None, w["e_mikL"],
                                device=local, rank=rank, nranks=world, **kw)
            uid = [fb.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            A.comm_init(uid[0])
            A.set_eps_objects(w["grid"].lg_prim, w["shapes"], w["pinds"], w["params"])
            return A
        return workloads.make_operator_from_objects(w, device=local, **kw)
    A = workloads.make_operator(w, device=local, rank=rank, nranks=world, **kw)
    if world > 1:
        uid = [fb.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        A.comm_init(uid[0])
    return A


def bytes_per_dof(A, w):
    """bytes one apply must move per DOF on this handle (x 16 + y 16 + diagonal mass 16 or 8 + off-diagonal streams on
    the (tile, plane) blocks that hold any: 32, or 16 when the tensor is symmetric and stored once)"""
    off = A.offdiag_fraction if w.get("full_eps") else 0.0
    sym = bool(w.get("full_eps") and A.offdiag_symmetric)
    return 32.0 + A.mass_bytes_per_dof + A.offdiag_bytes_per_dof * off, off, sym


def measure_config(name, w, local, rank, world, steps, kry_iters, gen, extra=None, **opkw):
    """device-resident apply throughput (+ min / median over batches), BiCGSTAB iterations/s for one workload"""
    import torch
    t0 = time.perf_counter()
    A = build_operator(w, local, rank, world, **opkw)
    n = A.n
    x = rand_vec(n, gen)
    y = torch.empty_like(x)
    A.bench_apply(x, y, warmup=10 if steps >= 50 else 3, iters=1)
    setup_s = time.perf_counter() - t0
    barrier(world)
    nb = 10
    per = max(1, steps // nb)
    batch = []
    for _ in range(nb):
        ms, _ = A.bench_apply(x, y, warmup=0, iters=per)
        batch.append(max_over_ranks(ms / per, world))
    ms_med, ms_min = statistics.median(batch), min(batch)
    bpd, off, sym = bytes_per_dof(A, w)
    if world > 1:      # slabs differ (a sphere / a pillar layer touches some of them only): report the mean over the ranks
        import torch.distributed as dist
        t = torch.tensor([bpd, off], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        bpd, off = float(t[0]) / world, float(t[1]) / world
    peak, _ = peaks()
    n_tot = 3 * w["N"][0] * w["N"][1] * w["N"][2]
    out = {"config": name, "grid": list(w["N"]), "dof": n_tot, "n_gpus": world, "applies_timed": nb * per,
           "ms_per_apply": ms_med, "ms_per_apply_min": ms_min, "gdof_s": n_tot / ms_med / 1e6,
           "gdof_s_best": n_tot / ms_min / 1e6, "bytes_per_dof": bpd, "offdiag_block_fraction": off,
           "offdiag_symmetric": sym, "hbm_frac": bpd * (n_tot / world) / (ms_med * 1e-3) / 1e9 / peak, "setup_s": setup_s}
    if kry_iters:
        try:
            b = rand_vec(n, gen)
            xs = torch.zeros_like(b)
            barrier(world)
            ms_k = max_over_ranks(A.bench_solve(b, xs, "bicgstab", warmup=2, iters=kry_iters), world)
            out["bicgstab_it_s"] = kry_iters / (ms_k * 1e-3)
            out["bicgstab_hbm_frac"] = (2 * bpd + 240) * (n_tot / world) * out["bicgstab_it_s"] / 1e9 / peak
            del b, xs
        except Exception as e:  # noqa: BLE001
            out["bicgstab_error"] = f"{type(e).__name__}: {e}"
    if extra:
        out.update(extra(A, x))
    A.close()
    del x, y
    torch.cuda.empty_cache()
    return out


def single_call_block(N, per, ngpu, steps, kry_iters):
    """e2e through the single-call boundary: MultiGpuOperator over `ngpu` devices driven by THIS process alone, on the same
    replicated-C2 grid as the headline; pinned full-grid host vectors; apply = mul!(y, A, x), solve = A \\ b with a fixed
    iteration count (b copied in, x copied out)."""
    import numpy as np
    import torch
    import workloads
    import maxwellfdm_jl_b200 as fb
    t0 = time.perf_counter()
    # one period of the replicated-C2 cross-section, tiled along z in the memory layout the C ABI takes (the Julia
    # column-major (Nx,Ny,Nz,3,3) array = C-order (3,3,Nz,Ny,Nx)): no 9 GB transposing copy at 8 GPUs
    w1 = workloads.c2_waveguide(per, 0, per[2])
    eps = np.tile(np.ascontiguousarray(w1["eps"].transpose(4, 3, 2, 1, 0)), (1, 1, ngpu, 1, 1))
    del w1
    w = workloads._common(N, 20.0, (False, False, False), ((10,) * 3, (10,) * 3))
    A = fb.MultiGpuOperator(w["N"], w["isbloch"], w["sdl_e"], w["sdl_m"], w["omega"], eps, None, w["e_mikL"], ngpu=ngpu)
    del w, eps
    n = A.n
    xh = torch.empty(n, dtype=torch.complex128).pin_memory()
    yh = torch.empty(n, dtype=torch.complex128).pin_memory()
    g = torch.Generator().manual_seed(7)
    xh.copy_(torch.randn(n, 2, dtype=torch.float64, generator=g).view(torch.complex128).reshape(-1))
    A.mul(yh.numpy(), xh.numpy())
    setup_s = time.perf_counter() - t0
    ts = []
    for _ in range(steps):
        t1 = time.perf_counter()
        A.mul(yh.numpy(), xh.numpy())
        ts.append(time.perf_counter() - t1)
    t_apply = statistics.median(ts)
    A.solve(xh.numpy(), rtol=1e-300, maxit=3, check_every=1 << 30, out=yh.numpy())
    yh.zero_()
    t1 = time.perf_counter()
    _, info = A.solve(xh.numpy(), rtol=1e-300, maxit=kry_iters, check_every=1 << 30, out=yh.numpy())
    t_solve = time.perf_counter() - t1
    A.close()
    return {"n_gpus": ngpu, "grid": list(N), "dof": n, "apply_gdof_s": n / t_apply / 1e9, "apply_ms": t_apply * 1e3,
            "apply_gdof_s_per_gpu": n / t_apply / 1e9 / ngpu, "h2d_bytes_per_apply": 16 * n + 32 * (ngpu - 1) * 3 * N[0] * N[1],
            "d2h_bytes_per_apply": 16 * n, "solve_iters": info["iters"], "solve_seconds": t_solve,
            "solve_iter_per_s": info["iters"] / t_solve, "solve_gdof_iter_per_s": n * info["iters"] / t_solve / 1e9,
            "setup_s": setup_s,
            "note": "ONE process / ONE handle over all GPUs (fdfd_multi_apply, fdfd_multi_solve): host threads = GPUs, halo "
                    "planes of an apply ride along with the H2D copies, Krylov halos and dots over NCCL"}


def parity_block(local, rank, world):
    """UNTIMED correctness check at this N (oracle = checker): a reduced copy of the C2 workload (full 3x3 eps with
    off-diagonals, PML) on world z-slabs - (a) y = A x on this rank's planes against the oracle's matrix-free numpy
    apply on the global grid, (b) five BiCGSTAB iterations against a textbook iteration on the oracle operator."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import workloads
    import maxwellfdm_jl_b200 as fb
    from oracle.matfree import MatFreeOperator
    N = (46, 40, 12 * world)
    wg = workloads.c2_waveguide(N, 0, N[2], npml=4)                     # global arrays for the oracle
    k0, k1 = fb.partition(N[2], world, rank)
    w = dict(wg, eps=np.ascontiguousarray(wg["eps"][:, :, k0:k1]), k0=k0, k1=k1)
    A = build_operator(w, local, rank, world)
    rng = np.random.default_rng(20261018)
    nz = N[0] * N[1] * 3
    xg = rng.standard_normal(nz * N[2]) + 1j * rng.standard_normal(nz * N[2])
    O = MatFreeOperator(0, wg["omega"], wg["eps"], None, wg["sdl_e"], wg["sdl_m"], (0, 0, 0), wg["isbloch"], wg["e_mikL"])
    yg = O.apply(xg)
    sl = slice(nz * k0, nz * k1)
    yl = (A @ torch.from_numpy(xg[sl].copy()).cuda()).cpu().numpy()
    num, den = np.linalg.norm(yl - yg[sl]) ** 2, np.linalg.norm(yg[sl]) ** 2
    # textbook BiCGSTAB (5 iterations, x0 = 0) on the oracle operator
    bg = O.apply(rng.standard_normal(xg.size) + 1j * rng.standard_normal(xg.size))
    xr = np.zeros_like(bg); r = bg.copy(); rh = r.copy(); rho = alpha = om = 1.0 + 0j; v = np.zeros_like(bg); p = np.zeros_like(bg)
    for _ in range(5):
        rho1 = np.vdot(rh, r); beta = (rho1 / rho) * (alpha / om); rho = rho1
        p = r + beta * (p - om * v); v = O.apply(p); alpha = rho / np.vdot(rh, v)
        s = r - alpha * v; t = O.apply(s); om = np.vdot(t, s) / np.vdot(t, t)
        xr = xr + alpha * p + om * s; r = s - om * t
    xs, _ = A.solve(torch.from_numpy(bg[sl].copy()).cuda(), rtol=1e-300, maxit=5, check_every=1)
    xs = xs.cpu().numpy()
    num2, den2 = np.linalg.norm(xs - xr[sl]) ** 2, np.linalg.norm(xr[sl]) ** 2
    A.close()
    t = torch.tensor([num, den, num2, den2], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t)
    num, den, num2, den2 = (float(v) for v in t.tolist())
    return {"grid": list(N), "n_slabs": world, "apply_rel_err": (num / den) ** 0.5, "traj_rel_err": (num2 / den2) ** 0.5,
            "checker": "oracle/matfree.py on the global grid (untimed)", "tolerance": {"apply": 1e-12, "trajectory": 1e-9}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--krylov-iters", type=int, default=200)   # fixed 200 iterations, set-up included
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-configuration block (N = 1)")
    ap.add_argument("--no-scale", action="store_true", help="skip the C5 weak / C4 strong scaling blocks")
    ap.add_argument("--no-single-call", action="store_true", help="skip the one-process / N-GPU boundary measurement")
    ap.add_argument("--diag", action="store_true", help="diagonal-eps variant of the workload (48 B/DOF)")
    ap.add_argument("--dense-off", action="store_true",
                    help="variant with non-zero off-diagonal eps in EVERY cell (dense full-tensor kernel path)")
    ap.add_argument("--n", type=int, nargs=3, default=None, help="override the per-GPU grid (debug)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import workloads
    import maxwellfdm_jl_b200 as fb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        cpu_group = dist.new_group(backend="gloo")     # host-side rendezvous that keeps the GPUs idle (single-call block)
    gen = torch.Generator(device="cuda").manual_seed(20261017 + rank)
    peak, peak_src = peaks()

    # ---- untimed parity check at this N (oracle as the checker) ---------------------------------------------
    try:
        parity = parity_block(local, rank, world)
    except Exception as e:  # noqa: BLE001   (reported, never hidden; the timed numbers below still stand on their own)
        parity = {"error": f"{type(e).__name__}: {e}"}

    per = tuple(args.n) if args.n else PER_GPU_N
    N = (per[0], per[1], per[2] * world)
    k0, k1 = fb.partition(N[2], world, rank)
    w = workloads.c2_waveguide(N, k0, k1, period_z=per[2])
    if args.diag:
        for v in range(3):
            for u in range(3):
                if u != v:
                    w["eps"][..., v, u] = 0
        w["full_eps"] = False
    if args.dense_off:
        workloads.make_dense_offdiag(w, seed=7 + rank)
    A = build_operator(w, local, rank, world)
    n_loc = A.n
    n_tot = 3 * N[0] * N[1] * N[2]
    x = rand_vec(n_loc, gen)
    y = torch.empty_like(x)
    torch.cuda.synchronize()

    # ---- device-resident operator throughput ------------------------------------------------
    A.bench_apply(x, y, warmup=max(args.warmup, 10), iters=1)
    sampler = ClockSampler(local).start_and_wait()
    barrier(world)
    l0 = A.launch_count
    ms_total, _ = A.bench_apply(x, y, warmup=0, iters=args.steps)
    barrier(world)
    launches = A.launch_count - l0
    clocks = sampler.summary()
    ms_step = max_over_ranks(ms_total, world) / args.steps
    gdofs = n_tot / (ms_step * 1e-3) / 1e9
    # BASELINE.md section 4 protocol: >= 100 further applies in 10 batches, median and min per apply
    batch = []
    for _ in range(10):
        ms_b, _ = A.bench_apply(x, y, warmup=0, iters=10)
        batch.append(max_over_ranks(ms_b / 10, world))
    ms_med, ms_min = statistics.median(batch), min(batch)

    # ---- halo exchange by itself (N > 1) ----------------------------------------------------------------------
    halo = None
    if world > 1:
        try:
            ms_h, nb = A.bench_halo(x, warmup=5, iters=50)
            us = max_over_ranks(ms_h / 50 * 1e3, world)
            nbm = int(max_over_ranks(nb, world))
            halo = {"bytes_sent_per_rank": nbm, "us": us, "gbs_per_direction": nbm / 2 / (us * 1e-6) / 1e9 if us > 0 else None,
                    "nvlink_frac": (nbm / 2 / (us * 1e-6) / 1e9 / NVLINK_PEER_GBS) if us > 0 else None,
                    "nvlink_peak_gbs": NVLINK_PEER_GBS, "share_of_apply": us / (ms_step * 1e3),
                    "data_plane": A.halo_data_plane,
                    "note": "the exchange of the two boundary planes timed alone. data_plane 'peer': copy-engine copies into "
                            "IPC-mapped neighbour buffers ordered by stream memory operations, hidden behind the interior "
                            "z-chunks of the apply kernel (in-kernel halo wait); 'nccl': grouped ncclSend/ncclRecv, serialised "
                            "in front of a plain apply. Inside BiCGSTAB either one starts early and hides behind the vector "
                            "kernels"}
        except Exception as e:  # noqa: BLE001
            halo = {"error": f"{type(e).__name__}: {e}"}

    # ---- end to end through the C ABI with host buffers ------------------------------------------
    xh = torch.empty(n_loc, dtype=torch.complex128).pin_memory()
    yh = torch.empty(n_loc, dtype=torch.complex128).pin_memory()
    xh.copy_(x.cpu())
    e2e_steps = max(3, min(args.steps, 5))
    A.mul(yh.numpy(), xh.numpy())
    barrier(world)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        A.mul(yh.numpy(), xh.numpy())
    barrier(world)
    e2e = n_tot / max_over_ranks((time.perf_counter() - t0) / e2e_steps, world) / 1e9

    # ---- Krylov iterations / s (2 applies + fused vector updates per BiCGSTAB iteration) ---------
    kry_vec = 256 if os.environ.get("FDFD_BICGSTAB_CLASSIC") else 240
    it_per_s = qmr_it_per_s = krylov_error = None
    e2e_solve = None
    try:
        b = rand_vec(n_loc, gen)
        xs = torch.zeros_like(b)
        barrier(world)
        ms_k = max_over_ranks(A.bench_solve(b, xs, "bicgstab", warmup=2, iters=args.krylov_iters), world)
        it_per_s = args.krylov_iters / (ms_k * 1e-3)
        xs.zero_()
        barrier(world)
        ms_q = max_over_ranks(A.bench_solve(b, xs, "qmr", warmup=2, iters=args.krylov_iters), world)
        qmr_it_per_s = args.krylov_iters / (ms_q * 1e-3)
        # the path's real end-to-end unit: A \ b with host buffers - b in, x out, fixed iteration count
        xh.copy_(b.cpu())
        A.solve(xh.numpy(), rtol=1e-300, maxit=3, check_every=1 << 30)
        barrier(world)
        t0 = time.perf_counter()
        _, info = A.solve(xh.numpy(), rtol=1e-300, maxit=args.krylov_iters, check_every=1 << 30)
        barrier(world)
        ts = max_over_ranks(time.perf_counter() - t0, world)
        e2e_solve = {"iters": info["iters"], "seconds": ts, "iter_per_s": info["iters"] / ts,
                     "gdof_iter_per_s": n_tot * info["iters"] / ts / 1e9,
                     "h2d_bytes": 16 * n_loc * world, "d2h_bytes": 16 * n_loc * world,
                     "note": "fdfd_solve(FDFD_HOST): b copied in, x copied out, BiCGSTAB without convergence exit"}
        del b, xs
    except Exception as e:  # noqa: BLE001
        krylov_error = f"{type(e).__name__}: {e}"
        print(f"bench.py: Krylov timing failed: {krylov_error}", file=sys.stderr)

    bpd, off_frac, off_sym = bytes_per_dof(A, w)
    mass_b = A.mass_bytes_per_dof
    A.close()
    del x, y, xh, yh
    torch.cuda.empty_cache()

    # ---- the other BASELINE configurations (N = 1) and the north_star scaling configurations (every N) ----------
    configs, scale_c5, scale_c4 = None, None, None

    def guarded(fn):
        try:
            return fn()
        except Exception as e:  # noqa: BLE001
            torch.cuda.empty_cache()
            return {"error": f"{type(e).__name__}: {e}"}

    if world == 1 and not args.no_configs:
        configs = []
        configs.append(guarded(lambda: measure_config("C1 vacuum box 40^3 + PML", workloads.c1_vacuum_box(), local, 0, 1,
                                                       300, 50, gen)))
        configs.append(guarded(lambda: measure_config("C3 PhC slab 256x256x128, Bloch x/y, PML z", workloads.c3_phc_slab(),
                                                       local, 0, 1, 100, 40, gen)))
        wd = workloads.c2_waveguide(PER_GPU_N)
        workloads.make_dense_offdiag(wd, seed=7)
        configs.append(guarded(lambda: measure_config("C2 grid, dense off-diagonal eps (fused full-tensor kernel)", wd, local,
                                                       0, 1, 100, 40, gen)))
        del wd
        configs.append(guarded(lambda: measure_config("C2 grid, HH formulation A = Ce eps^-1 Cm - w^2 mu (model.jl:238-240)",
                                                       workloads.c2_hh(), local, 0, 1, 100, 40, gen, ft="H")))
    if not args.no_scale and not args.n:
        # C4: 512^3 scatterer from objects (Kottke-smoothed on the device), STRONG scaling over the N slabs
        scale_c4 = guarded(lambda: measure_config("C4 dielectric sphere 512^3 (strong scaling: 512/N planes per GPU)",
                                                  workloads.c4_objects(), local, rank, world, 20, 10, gen))
        # C5: metalens, WEAK scaling unit 1024 x 1024 x 96 planes per GPU (N = 8: 1024 x 1024 x 768)
        scale_c5 = guarded(lambda: measure_config("C5 metalens 1024x1024x(96 N) (weak scaling: 96 planes per GPU)",
                                                  workloads.c5_objects(N=(1024, 1024, 96 * world)), local, rank, world,
                                                  20, 10, gen))

    # ---- the drop-in boundary itself: ONE process, ONE call, `world` GPUs (fdfd_multi_*) ------------------
    # Rank 0 alone drives every GPU of the job through a single handle - full-grid pinned host vectors in and out, the
    # library cuts the z-slabs (one host thread per device) - while the other ranks wait at the barrier below.
    single_call = None
    if not args.n and not args.no_single_call:
        # the other ranks must leave their GPUs IDLE while rank 0 drives them: an NCCL barrier would park a spinning kernel
        # of another process on every GPU (measured: 319 instead of 660 it/s at 2 GPUs), so they wait on a CPU (gloo) barrier
        barrier(world)
        if rank == 0:
            single_call = guarded(lambda: single_call_block(N, per, world, e2e_steps, args.krylov_iters))
        if world > 1:
            dist.barrier(group=cpu_group)

    if rank == 0:
        achieved = bpd * (n_tot / world) / (ms_step * 1e-3) / 1e9      # per GPU
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get("dense" if args.dense_off else ("c2" if w["full_eps"] else "diag"))
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": gdofs, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "ms_per_step_median": ms_med, "ms_per_step_min": ms_min,
            "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "c128 (complex fp64)", "data": "synthetic",
            "config": {"workload": w["name"] + (" [diagonal-eps variant]" if args.diag else "") +
                       (" [dense off-diagonal variant]" if args.dense_off else ""),
                       "offdiag_block_fraction": off_frac, "offdiag_symmetric": off_sym,
                       "diagonal_mass_bytes_per_dof": mass_b,
                       "grid": list(N), "per_gpu_grid": list(per), "dof": n_tot, "parallelism": f"z-slab x{world}",
                       "l2": "inputs (x, y, eps: ~1 GB per GPU) larger than the 126 MB L2; no flush needed",
                       "bytes_per_dof": bpd},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "kernel": "apply_rowpair_kernel (+ offdiag_march_kernel on the flagged blocks when off-diagonal "
                                   "eps is sparse); traffic from the ncu --set full capture in profiles/",
                         "bytes_per_dof": bpd,
                         "bytes_model": "x 16 + y 16 + diagonal mass %g (real entries travel as doubles) + off-diagonal "
                                        "streams on the flagged blocks" % mass_b},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 16 * n_loc * world,
                    "d2h_bytes_per_step": 16 * n_loc * world, "steps": e2e_steps,
                    "note": "fdfd_apply(FDFD_HOST) with pinned host buffers"},
            "e2e_solve": e2e_solve,
            "gpu_launches": launches, "clocks": clocks, "parity": parity,
            "krylov": {"method": "bicgstab", "iters": args.krylov_iters, "iter_per_s": it_per_s,
                       "gdof_iter_per_s": None if it_per_s is None else n_tot * it_per_s / 1e9,
                       # 15 vector passes of 16 B (s: 3, x/r update with both dots: 7, p with the next sigma: 5)
                       "bytes_per_dof_model": 2 * bpd + kry_vec,
                       "hbm_frac": None if it_per_s is None else (2 * bpd + kry_vec) * (n_tot / world) * it_per_s / 1e9 / peak,
                       "error": krylov_error,
                       "qmr_iter_per_s": qmr_it_per_s, "qmr_bytes_per_dof_model": 2 * bpd + 304},
        }
        if single_call is not None:
            line["e2e_single_call"] = single_call
        if halo is not None:
            line["halo"] = halo
        if configs is not None:
            line["configs"] = configs
        if scale_c4 is not None:
            line["scale_c4"] = scale_c4
        if scale_c5 is not None:
            line["scale_c5"] = scale_c5
        if world == 1 and not args.no_cpu:
            try:
                line["cpu_baseline"] = cpu_baseline_port()
            except Exception as e:  # the baseline is a report, never a reason to lose the GPU numbers
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "port",
                                        "sample": f"failed: {e}"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
