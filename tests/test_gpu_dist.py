"""Real multi-GPU run of the z-slab path (NCCL halo exchange + allreduce): needs >= 2 GPUs, else skipped.
The same script runs under `gpurun --gpus N` (scripts/dist_check.py)."""
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
