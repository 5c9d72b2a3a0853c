"""Needs >= 2 GPUs (gpurun --gpus 2): DDP / ZeRO-1 over NCCL vs a single-process run with the same total batch."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_ddp_and_zero1_match_single_process():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", str(ROOT / "scripts" / "dev" / "dp_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "ddp:" in res.stdout and "zero1:" in res.stdout and "roberta zero1 vs ddp" in res.stdout and "FAIL" not in res.stdout
