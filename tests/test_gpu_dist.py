"""Real multi-GPU runs of the z-slab path (NCCL halo exchange + allreduce), one process per GPU (scripts/dist_check.py) and
ONE process over N GPUs (fdfd_multi_*, scripts/multi_check.py): need >= 2 GPUs, else skipped - except the single-call
handle with ngpu = 1, which runs everywhere.  The same scripts run under `gpurun --gpus N`."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slabs_over_nccl():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 2 if n < 4 else 4
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29631",
                        os.path.join(ROOT, "scripts", "dist_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "DIST_CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_single_call_handle_over_the_gpus_of_the_box():
    """fdfd_multi_* (one host thread per device, full-grid host arrays in, slabs cut by the library) against the oracle:
    ngpu = 1 always, 2 / 4 when the box has them.  Seam: one value from one process, model.jl:209-246."""
    import torch
    n = torch.cuda.device_count()
    if n < 1:
        pytest.skip("needs a GPU")
    want = [str(g) for g in (1, 2, 4) if g <= n]
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "multi_check.py"), *want], capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0 and "MULTI_CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
