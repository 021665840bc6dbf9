#!/usr/bin/env python
"""bench.py - FDFD operator throughput (GDOF/s) on B200, BASELINE.json metric.

A "step" is ONE application y = A x of the matrix-free operator A = curl mu^-1 curl - w^2 eps on the
workload of BASELINE.json configs[1] (200^3 Si strip waveguide, full 3x3 eps, 10-cell PML); with N GPUs
the grid is 200 x 200 x (200 N) split into N z-slabs (weak scaling, one NCCL halo exchange per apply).
    value     GDOF/s, x / eps resident in HBM, CUDA events on the library's stream, max over ranks
    e2e       the same metric through the C-ABI call with HOST (pinned) buffers: H2D of x and D2H of y
              inside the timed region
    roofline  algorithmic bytes (80 B/DOF full-tensor, 48 B/DOF diagonal) / apply time
              against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
    cpu_baseline   the oracle's Julia-style single-thread CSC mul! on a bounded sample of the workload
--impl reference times the reference's CPU path stand-in (CSC assembled by the oracle restatement,
product on all host cores) - the reference itself is Julia and cannot run here.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fdfd_operator_apply_throughput"
UNIT = "GDOF/s"
PER_GPU_N = (200, 200, 200)
SAMPLE_PLANES = 8


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled through NVML every ~2 ms while the timed region runs
    (an nvidia-smi subprocess is too slow for a region of a few tens of ms)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.err = index, [], False, None
        self.max_mhz = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # honour CUDA_VISIBLE_DEVICES-style remapping by matching the torch device's UUID when possible
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            try:
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
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self.stop_flag:
                self.samples.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(get_reasons(h))))
                time.sleep(0.002)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable: " + str(self.err)]}
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                "hw_power_brake_slowdown": 0x80}
        allr = 0
        for _, r in self.samples:
            allr |= r
        return {"sm_mhz": statistics.median(c for c, _ in self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": [n for n, b in bits.items() if allr & b], "samples": len(self.samples)}


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
    t0 = time.perf_counter()
    for _ in range(args.steps):
        R.mul(x, y)
    dt = (time.perf_counter() - t0) / args.steps
    cores = cb.num_threads()
    val = n / dt / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "c128 (complex fp64)", "data": "synthetic",
            "config": {"workload": "C2 Si strip waveguide 200x200x200, full 3x3 eps, 10-cell PML",
                       "sample": desc, "l2": "matrix stream larger than LLC"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": desc + "; OpenMP CSR product of the oracle-assembled matrix (stand-in for the "
                                              "Julia SparseMatrixCSC mul!, which cannot run here: no julia binary)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--krylov-iters", type=int, default=200)   # fixed 200 iterations, set-up included
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--diag", action="store_true", help="diagonal-eps variant of the workload (48 B/DOF)")
    ap.add_argument("--dense-off", action="store_true",
                    help="variant with non-zero off-diagonal eps in EVERY cell (80 B/DOF: dense full-tensor kernel path)")
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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

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
        rng = np.random.default_rng(7 + rank)
        for (v, u) in ((0, 1), (0, 2), (1, 2)):
            pert = 0.05 * (rng.random(w["eps"].shape[:3]) - 0.5)
            w["eps"][..., v, u] = pert
            w["eps"][..., u, v] = pert
    A = workloads.make_operator(w, device=local, rank=rank, nranks=world)
    if world > 1:
        uid = [fb.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        A.comm_init(uid[0])
    n_loc = A.n
    n_tot = 3 * N[0] * N[1] * N[2]
    g = torch.Generator(device="cuda").manual_seed(20261017 + rank)
    x = torch.randn(n_loc, 2, device="cuda", dtype=torch.float64, generator=g).view(torch.complex128).reshape(-1)
    y = torch.empty_like(x)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident operator throughput ------------------------------------------------
    A.bench_apply(x, y, warmup=args.warmup, iters=1)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    l0 = A.launch_count
    ms_total, _ = A.bench_apply(x, y, warmup=0, iters=args.steps)
    barrier()
    launches = A.launch_count - l0
    clocks = sampler.summary()
    t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    gdofs = n_tot / (ms_step * 1e-3) / 1e9

    # ---- end to end through the C ABI with host buffers ------------------------------------------
    xh = torch.empty(n_loc, dtype=torch.complex128).pin_memory()
    yh = torch.empty(n_loc, dtype=torch.complex128).pin_memory()
    xh.copy_(x.cpu())
    e2e_steps = max(3, min(args.steps, 5))
    A.mul(yh.numpy(), xh.numpy())
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        A.mul(yh.numpy(), xh.numpy())
    barrier()
    te = torch.tensor([(time.perf_counter() - t0) / e2e_steps], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e = n_tot / float(te.item()) / 1e9

    # ---- Krylov iterations / s (2 applies + fused vector updates per BiCGSTAB iteration) ---------
    kry_vec = 256 if os.environ.get("FDFD_BICGSTAB_CLASSIC") else 240
    # (a failure here must not lose the operator numbers measured above: it is reported in the line instead)
    it_per_s = qmr_it_per_s = krylov_error = None
    try:
        b = torch.randn(n_loc, 2, device="cuda", dtype=torch.float64, generator=g).view(torch.complex128).reshape(-1)
        xs = torch.zeros_like(b)
        barrier()
        ms_k = A.bench_solve(b, xs, "bicgstab", warmup=2, iters=args.krylov_iters)
        tk = torch.tensor([ms_k], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tk, op=dist.ReduceOp.MAX)
        it_per_s = args.krylov_iters / (float(tk.item()) * 1e-3)
        xs.zero_()
        barrier()
        ms_q = A.bench_solve(b, xs, "qmr", warmup=2, iters=args.krylov_iters)
        tq = torch.tensor([ms_q], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tq, op=dist.ReduceOp.MAX)
        qmr_it_per_s = args.krylov_iters / (float(tq.item()) * 1e-3)
        del b, xs
    except Exception as e:  # noqa: BLE001
        krylov_error = f"{type(e).__name__}: {e}"
        print(f"bench.py: Krylov timing failed: {krylov_error}", file=sys.stderr)

    if rank == 0:
        peak, peak_src = peaks()
        # bytes an apply must move per DOF: x 16 + y 16 + eps_diag 16, + 32 for the six off-diagonal entries on the
        # (tile, plane) blocks that hold any (the kernel skips empty blocks; dense off-diagonals -> 80)
        # a pointwise symmetric tensor is stored once (three arrays): 16 instead of 32
        off_frac = A.offdiag_fraction if w["full_eps"] else 0.0
        off_sym = bool(w["full_eps"] and A.offdiag_symmetric)
        bpd = 48 + (16 if off_sym else 32) * off_frac
        achieved = bpd * (n_tot / world) / (ms_step * 1e-3) / 1e9      # per GPU
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get("dense" if args.dense_off else ("c2" if w["full_eps"] else "diag"))
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": gdofs, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "c128 (complex fp64)", "data": "synthetic",
            "config": {"workload": w["name"] + (" [diagonal-eps variant]" if args.diag else "") +
                       (" [dense off-diagonal variant]" if args.dense_off else ""),
                       "offdiag_block_fraction": off_frac, "offdiag_symmetric": off_sym,
                       "bytes_per_dof_if_dense": (64 if off_sym else 80) if w["full_eps"] else 48,
                       "grid": list(N), "per_gpu_grid": list(per), "dof": n_tot, "parallelism": f"z-slab x{world}",
                       "l2": "inputs (x, y, eps: > 1 GB per GPU) larger than the 126 MB L2; no flush needed",
                       "bytes_per_dof": bpd},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "kernel": "apply_tiled_kernel (+ offdiag_correction_kernel on flagged blocks when off-diagonal "
                                   "eps is sparse); traffic from the ncu --set full capture in profiles/",
                         "bytes_per_dof": bpd},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 16 * n_loc * world,
                    "d2h_bytes_per_step": 16 * n_loc * world, "steps": e2e_steps,
                    "note": "fdfd_apply(FDFD_HOST) with pinned host buffers"},
            "gpu_launches": launches, "clocks": clocks,
            "krylov": {"method": "bicgstab", "iters": args.krylov_iters, "iter_per_s": it_per_s,
                       # 15 vector passes of 16 B (s: 3, x/r update with both dots: 7, p with the next sigma: 5);
                       # 16 with FDFD_BICGSTAB_CLASSIC (separate (rhat, v) pass)
                       "bytes_per_dof_model": 2 * bpd + kry_vec,
                       "hbm_frac": None if it_per_s is None else (2 * bpd + kry_vec) * (n_tot / world) * it_per_s / 1e9 / peak,
                       "error": krylov_error,
                       "qmr_iter_per_s": qmr_it_per_s, "qmr_bytes_per_dof_model": 2 * bpd + 304},
        }
        if world == 1 and not args.no_cpu:
            try:
                line["cpu_baseline"] = cpu_baseline_port()
            except Exception as e:  # the baseline is a report, never a reason to lose the GPU numbers
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "port",
                                        "sample": f"failed: {e}"}
        print(json.dumps(line))
    A.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
