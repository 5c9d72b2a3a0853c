"""find_max_mbs_pow2 — mirrors src/benchmarking/max_batch_size.py:11-25 (power-of-two search, OOM is the signal)."""
from __future__ import annotations

import logging

import torch.cuda

from .step_time import benchmark_acc_optim_times
from .utils import ManualTrainer

logger = logging.getLogger("academic-pretraining")


def find_max_mbs_pow2(trainer: ManualTrainer, limit: int) -> int:
    mbs = 1
    while mbs <= limit:
        logger.info(f"Running 1 training step with MBS = {mbs} ...")
        try:
            benchmark_acc_optim_times(trainer=trainer, micro_batch_size=mbs, training_steps=1, accumulations=1)
        except torch.cuda.OutOfMemoryError:
            trainer.model.zero_grad()
            break
        mbs *= 2
    return mbs // 2
