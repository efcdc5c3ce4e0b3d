"""Two-GPU data-parallel training through the package's kernels (examples/train_ddp.py, the upstream recipe's shape: 2-GPU DDP,
reference website/src/pages/[lang]/reprod/index.astro:238-252).  Skipped on a box with one GPU."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_ddp_training_step(built_lib):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "examples", "train_ddp.py"), "--steps", "12", "--frames", "8"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["world_size"] == 2 and line["params_identical_across_ranks"] and line["grad_finite"]
    assert line["loss_last"] < line["loss_first"]
    assert line["kernel_launches_rank0"] >= 3 * 12          # projection, training forward, backward per step
