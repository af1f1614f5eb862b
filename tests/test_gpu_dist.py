"""GPU, >= 2 devices: the gathered head (compat strips AND the fused kernel) + bucketed bf16 gradient all-reduce reproduce
the single-process global-batch step and the CPU oracle (tools/dist_check.py), at W = 2 and at every GPU of the box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, batch, layers, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "dist_check.py"), "--batch", str(batch),
           "--layers", str(layers)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "DIST_CHECK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_two_rank_step_matches_global_batch(cuda_dev):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, 32, 2, 29611)


def test_all_rank_step_matches_global_batch(cuda_dev):
    import torch
    n = torch.cuda.device_count()
    if n < 4:
        pytest.skip("needs >= 4 GPUs")
    _run(min(n, 8), 64, 3, 29613)
