"""Multi-GPU (NCCL) parity of the CFG-split / Ulysses path: launches tests/multi_rank_check.py under torchrun on every GPU
count the box offers (2, 4, 8).  Skipped on a single-GPU box, where tests/test_gpu_parallel.py covers the same data path
with virtual ranks."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n", [2, 4, 8])
def test_nccl_ranks_match_single_gpu(n):
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + n), os.path.join(ROOT, "tests", "multi_rank_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(r.stdout[-4000:], r.stderr[-4000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
