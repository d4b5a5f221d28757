"""The end-to-end denoise-loop example (branch + backbone + fused step end, chained windows with the ID-resample processor
and a merged LoRA) runs and stays finite; on a multi-GPU box the sharded run must reproduce the single-GPU latents bit for
bit (same SHA-256)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARGS = ["--steps", "3", "--layers", "2", "--windows", "2", "--resample", "--lora-rank", "32"]


def _run(prefix, args=None):
    r = subprocess.run(prefix + [os.path.join(ROOT, "examples", "inpaint_loop.py")] + (args or ARGS), capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_loop_example_single_and_multi_gpu_agree():
    one = _run([sys.executable])
    assert one["finite"] and one["ranks_agree"] and one["windows"] == 2
    n = min(torch.cuda.device_count(), 4)
    if n >= 4:
        many = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                     "--master-port", "29541"])
        assert many["ranks_agree"] and many["latents_sha256_16"] == one["latents_sha256_16"]


def test_loop_example_with_and_without_cuda_graphs_agree():
    """Chained windows through the module forwards: the second window reads the first window's hidden-state list, which in graph
    mode is a set of views of the first graph's arena (videopainter_b200/graphs.py)."""
    args = ["--steps", "4", "--layers", "2", "--windows", "3", "--resample"]
    eager = _run([sys.executable], args + ["--graphs", "0"])
    graphed = _run([sys.executable], args + ["--graphs", "1"])
    assert eager["finite"] and graphed["finite"] and graphed["cuda_graphs"] and not eager["cuda_graphs"]
    assert graphed["latents_sha256_16"] == eager["latents_sha256_16"]
