"""Needs >= 2 GPUs (gpurun --gpus 2): DDP / ZeRO-1 / ZeRO-2 over NCCL vs a single-process run with the same total batch.
The round's hardware logs of scripts/dev/dp_check.py (2 and 8 GPUs) are tracked under profiles/."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_ddp_zero1_zero2_match_single_process():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", str(ROOT / "scripts" / "dev" / "dp_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert all(k in res.stdout for k in ("ddp:", "zero1:", "zero2:", "zero2 vs ddp", "roberta zero1 vs ddp")) and "FAIL" not in res.stdout, res.stdout[-3000:]
