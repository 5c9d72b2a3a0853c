"""benchmark_acc_optim_times / estimate_step_time — mirror src/benchmarking/step_time.py:33-97.

One deliberate difference: the reference times with host `time.perf_counter()` and never synchronises the device
(src/benchmarking/step_time.py:14-18,55-61), which mis-attributes asynchronous GPU work. Here both
phases are DEVICE-timed with CUDA events on the launching stream."""
from __future__ import annotations

import gc
import logging
from contextlib import contextmanager

import torch

from .utils import ManualTrainer

logger = logging.getLogger("academic-pretraining")


class _EventTimer:
    def __init__(self):
        self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def seconds(self) -> float:
        self.e1.synchronize()
        return self.e0.elapsed_time(self.e1) / 1e3


@contextmanager
def perf_timer():
    t = _EventTimer()
    t.e0.record()
    yield t.seconds
    t.e1.record()


@contextmanager
def get_train_dataloader(trainer: ManualTrainer, micro_batch_size: int):
    original = trainer.args.per_device_train_batch_size
    try:
        trainer.args.per_device_train_batch_size = micro_batch_size
        yield iter(trainer.get_train_dataloader())
    finally:
        trainer.args.per_device_train_batch_size = original


def benchmark_acc_optim_times(trainer: ManualTrainer, micro_batch_size: int, training_steps: int = 1, accumulations: int = 1,
                              warmup: bool = False) -> tuple[float, float]:
    gc.collect()
    torch.cuda.empty_cache()
    acc_timers, opt_timers = [], []
    if warmup:
        training_steps += 1
    model = trainer.model_wrapped
    with get_train_dataloader(trainer, micro_batch_size) as train_dataloader:
        for _ in range(training_steps):
            for _ in range(accumulations):
                inputs = next(train_dataloader)
                with perf_timer() as t:
                    trainer.manual_training_step(model, inputs)
                acc_timers.append(t)
            with perf_timer() as t:
                trainer.manual_optimization_step(model)
            opt_timers.append(t)
    torch.cuda.synchronize()
    accumulation_times = [t() for t in acc_timers]
    optimization_times = [t() for t in opt_timers]
    if warmup:
        accumulation_times = accumulation_times[accumulations:]
        optimization_times = optimization_times[1:]
    logger.info(f"Accumulation times: {accumulation_times}")
    logger.info(f"Optimization times: {optimization_times}")
    return sum(accumulation_times) / len(accumulation_times), sum(optimization_times) / len(optimization_times)


def estimate_step_time(trainer: ManualTrainer, micro_batch_size: int, target_micro_batch_size: int,
                       num_benchmarking_steps: int) -> float:
    """step = acc_time * (target_mbs / mbs) + optim_time (src/benchmarking/step_time.py:75-97)."""
    accumulation_steps = target_micro_batch_size // micro_batch_size
    logger.info(f"Estimating step time for MBS = {micro_batch_size}, ACC = {accumulation_steps}")
    mean_acc_time, mean_optim_time = benchmark_acc_optim_times(trainer, micro_batch_size, training_steps=num_benchmarking_steps,
                                                               accumulations=1, warmup=True)
    return mean_acc_time * accumulation_steps + mean_optim_time


def compute_training_days(training_steps: int, step_time: float) -> float:
    """experiments/training_time_empirical.py:133-138."""
    return training_steps * step_time / 86400.0
