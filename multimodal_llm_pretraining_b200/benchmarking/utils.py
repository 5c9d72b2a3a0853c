"""ManualTrainer — mirrors src/benchmarking/utils.py:40-80: the harness object that exposes one micro-batch of
fwd+bwd (`manual_training_step`) and one optimizer step (`manual_optimization_step`) separately so they can be timed.

In the reference this is a re-classed transformers.Trainer (prepared by running trainer.train() up to the first
on_step_begin); here it is built directly on the B200 step engine: same two entry points, same argument order, plus
the pieces of HF Trainer state the harness reads (`args`, `model_wrapped`, `optimizer`, `lr_scheduler`,
`get_train_dataloader()`)."""
from __future__ import annotations

import os
from types import SimpleNamespace
from typing import Any

import torch
import torch.distributed as dist

from ..engine import TrainEngine
from ..optim import fused_optimizer_class, get_scheduler
from .data import ShardedBatchIterator


def strategy_for(world_size: int, zero_stage: str = "0", fsdp_sharding: str = "no_shard") -> str:
    """Sharding selection of the reference (src/train.py:126-181) -> engine strategy. DeepSpeed stage 2 and FSDP
    shard_grad_op are the same algorithm here (optimizer state + gradients sharded, parameters replicated)."""
    if world_size == 1:
        return "none"
    if zero_stage == "1":
        return "zero1"
    if zero_stage == "2" or fsdp_sharding == "shard_grad_op":
        return "zero2"
    if zero_stage == "0" and fsdp_sharding == "no_shard":
        return "ddp"
    if (zero_stage == "3" or fsdp_sharding == "full_shard") and os.environ.get("B200_EXPERIMENTAL_ZERO3"):
        return "zero3"  # engine.py: validated over gloo on CPU tensors only (tests/test_host_schedule_cpu.py), hence opt-in
    raise NotImplementedError(f"zero_stage={zero_stage!r} fsdp_sharding={fsdp_sharding!r}: parameter sharding (ZeRO-3 / FSDP full_shard) "
                              "is built in the engine (strategy 'zero3') but has not run on hardware yet — set B200_EXPERIMENTAL_ZERO3=1 "
                              "to select it; hybrid sharding and offload are outside this build's scope (SURVEY.md §2.3, §8f rank 3)")


class ManualTrainer:
    def __init__(self, model, args: dict[str, Any], train_dataset, optimizer_cls_and_kwargs, scheduler_type="linear",
                 zero_stage: str = "0", fsdp_sharding: str = "no_shard", device: torch.device | None = None, seed: int = 0):
        self.args = SimpleNamespace(**args)
        self.args.train_batch_size = self.args.per_device_train_batch_size
        self.train_dataset = train_dataset
        self.world_size = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = device
        if not hasattr(model, "flat"):
            raise NotImplementedError(
                "ManualTrainer drives B200 modules (build_model(use_custom_kernels=True)); the naive HF module of "
                "build_model(use_custom_kernels=False) is the parity oracle and the CPU baseline of bench.py --impl reference, "
                "it has no GPU training path in this build (no accelerate / HF Trainer in the image)")
        # precision of the reference's TrainingArguments (src/train.py:113-114): fp16 -> IEEE-half kernels + dynamic loss scaling
        if args.get("fp16"):
            model.set_compute_dtype(torch.float16)
        self.model = model.to(device).train()
        self.model_wrapped = self.model
        if self.args.gradient_checkpointing:
            self.model.gradient_checkpointing_enable()
        # HF Trainer.create_optimizer: two groups (decay / no-decay: biases and LayerNorm weights), both carrying
        # args.weight_decay (default 0.0), which overrides the kwarg (HF:trainer.py:1157-1198)
        opt_cls, opt_kwargs = optimizer_cls_and_kwargs
        wd = float(args.get("weight_decay", 0.0))
        decay = [p for n, p in self.model.named_parameters() if p.dim() >= 2]
        no_decay = [p for n, p in self.model.named_parameters() if p.dim() < 2]
        kw = {k: v for k, v in opt_kwargs.items() if k != "weight_decay"}
        opt_cls = fused_optimizer_class(opt_cls)  # torch.optim.Adam / AdamW -> the fused B200 classes (same signature and rule)
        self.optimizer = opt_cls([{"params": decay, "weight_decay": wd}, {"params": no_decay, "weight_decay": 0.0}], **kw)
        self.lr_scheduler = get_scheduler(self.args.lr_scheduler_type, self.optimizer, self.args.warmup_steps,
                                          self.args.max_steps, self.args.lr_scheduler_kwargs)
        strategy = strategy_for(self.world_size, zero_stage, fsdp_sharding)
        self.engine = TrainEngine(self.model, self.optimizer, self.lr_scheduler, max_grad_norm=self.args.max_grad_norm,
                                  gradient_accumulation_steps=self.args.gradient_accumulation_steps, strategy=strategy)
        self.seed = seed

    # -- the two timed calls (src/benchmarking/utils.py:61-80)
    def manual_training_step(self, model, inputs):
        inputs = {k: v.to(self.device, non_blocking=True) for k, v in inputs.items()}
        return self.engine.manual_training_step(inputs)

    def manual_optimization_step(self, model):
        return self.engine.manual_optimization_step()

    def get_train_dataloader(self):
        return ShardedBatchIterator(self.train_dataset, self.args.per_device_train_batch_size, self.world_size, self.rank,
                                    shuffle_seed=self.seed)
