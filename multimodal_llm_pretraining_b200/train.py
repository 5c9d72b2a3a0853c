"""TrainingClass — mirrors src/train.py:16-215 of the reference for the in-scope strategies.

`_to_huggingface_args_dict()` reproduces the reference's HF TrainingArguments dict key for key (README.md:78-123 of the
reference shows the expected output), including the FSDP option list and the DeepSpeed JSON for every ZeRO stage, so
`scripts/to_training_arguments.py` keeps emitting the same artefact. `build_trainer()` returns the B200 step engine
(multimodal_llm_pretraining_b200.benchmarking.utils.ManualTrainer) instead of transformers.Trainer. What runs on this
build: no sharding (DDP), zero_stage "1", zero_stage "2" / fsdp "shard_grad_op" (gradient sharding), bf16 or fp16 (+ dynamic
loss scaling), with or without activation checkpointing. Parameter sharding (zero_stage "3" / fsdp "full_shard") exists in the engine
(strategy "zero3") but has only been validated over gloo on CPU tensors, so it is selectable here only with
B200_EXPERIMENTAL_ZERO3=1; hybrid sharding and offload are out of scope.
"""
from __future__ import annotations

import gc
import os
from dataclasses import dataclass, field
from typing import Any, Literal

import torch
import torch.optim
from torch import nn
from torch.utils.data import Dataset


def _is_adam(opt_cls) -> bool:
    from .optim import B200Adam, B200AdamW

    return opt_cls in (torch.optim.Adam, torch.optim.AdamW, B200Adam, B200AdamW)


def _is_adamw(opt_cls) -> bool:
    from .optim import B200AdamW

    return opt_cls in (torch.optim.AdamW, B200AdamW)


@dataclass
class TrainingClass:
    num_training_steps: int
    micro_batch_size: int
    gradient_accumulation_steps: int
    gradient_checkpointing: bool = False
    bf16: bool = False
    fp16: bool = False
    tf32: bool = False
    compile: bool = False

    optimizer: type[torch.optim.Optimizer] = torch.optim.AdamW
    optimizer_kwargs: dict[str, Any] = field(default_factory=dict)
    scheduler_type: Any = "linear"
    scheduler_kwargs: dict[str, Any] = field(default_factory=dict)

    FsdpShardingT = Literal["no_shard", "shard_grad_op", "full_shard", "hybrid_shard_zero2", "hybrid_shard"]
    fsdp_sharding: FsdpShardingT = "no_shard"
    fsdp_layers_to_wrap: list[str] = field(default_factory=list)
    fsdp_offload: bool = False

    ZeroStageT = Literal["0", "1", "2", "3", "3++"]
    zero_stage: ZeroStageT = "0"
    zero_offload_optimizer: bool = False
    zero_offload_params: bool = False

    max_grad_norm: float = 1.0
    hf_training_args_overrides: dict[str, Any] = field(default_factory=dict)

    def is_valid(self) -> bool:  # src/train.py:45-55
        return not (
            self.num_training_steps <= 0
            or self.micro_batch_size <= 0
            or self.gradient_accumulation_steps <= 0
            or (self.bf16 and self.fp16)
            or (self.fsdp_sharding != "no_shard" and self.zero_stage != "0")
            or (self.fsdp_offload and self.fsdp_sharding == "no_shard")
            or (self.zero_offload_optimizer and self.zero_stage == "0")
            or (self.zero_offload_params and self.zero_stage not in ["3", "3++"])
        )

    def runs_on_b200_engine(self) -> bool:
        if self.zero_offload_optimizer or self.zero_offload_params or self.fsdp_offload:
            return False
        if self.fsdp_sharding in ("no_shard", "shard_grad_op") and self.zero_stage in ("0", "1", "2"):
            return True
        # parameter sharding: engine strategy "zero3", not yet run on hardware -> opt-in (benchmarking/utils.py: strategy_for)
        return bool(os.environ.get("B200_EXPERIMENTAL_ZERO3")) and (
            (self.zero_stage == "3" and self.fsdp_sharding == "no_shard") or (self.fsdp_sharding == "full_shard" and self.zero_stage == "0"))

    def build_trainer(self, model: nn.Module, train_dataset: Dataset, hf_training_args_overrides: dict[str, Any] = {},
                      hf_trainer_kwargs_overrides: dict[str, Any] = {}):
        """src/train.py:57-89. Returns a ManualTrainer (B200 step engine) with the optimizer hand-off of the reference:
        the model class's (optimizer, kwargs) tuple, weight decay overridden to TrainingArguments.weight_decay = 0.0
        (HF:trainer.py:1157-1170)."""
        from .benchmarking.utils import ManualTrainer

        gc.collect()
        if torch.cuda.is_available():
            torch.cuda.empty_cache()
        if not self.runs_on_b200_engine():
            raise NotImplementedError(
                f"sharding fsdp={self.fsdp_sharding!r} zero={self.zero_stage!r} offload is outside this build's scope "
                "(DDP, ZeRO-1, ZeRO-2 / FSDP shard_grad_op; SURVEY.md §2.3)")
        args = self._to_huggingface_args_dict(**hf_training_args_overrides)
        return ManualTrainer(model=model, args=args, train_dataset=train_dataset, optimizer_cls_and_kwargs=(self.optimizer, dict(self.optimizer_kwargs)),
                             scheduler_type=self.scheduler_type, zero_stage=self.zero_stage, fsdp_sharding=self.fsdp_sharding, **hf_trainer_kwargs_overrides)

    def to_huggingface_args(self, **hf_training_args_overrides):
        from transformers import TrainingArguments  # needs accelerate; not available in every image

        return TrainingArguments(**self._to_huggingface_args_dict(**hf_training_args_overrides))

    def _to_huggingface_args_dict(self, **hf_training_args_overrides) -> dict:  # src/train.py:94-124
        fsdp_options, fsdp_config = self._build_fsdp_config()
        ds_config = self._build_deepspeed_config()
        scheduler_kwargs = dict(self.scheduler_kwargs)
        scheduler_warmup_steps = scheduler_kwargs.pop("num_warmup_steps", 0)
        return dict(
            max_steps=self.num_training_steps,
            per_device_train_batch_size=self.micro_batch_size,
            gradient_accumulation_steps=self.gradient_accumulation_steps,
            lr_scheduler_type=getattr(self.scheduler_type, "value", self.scheduler_type),
            lr_scheduler_kwargs=scheduler_kwargs,
            warmup_steps=scheduler_warmup_steps,
            gradient_checkpointing=self.gradient_checkpointing,
            bf16=self.bf16,
            fp16=self.fp16,
            tf32=self.tf32,
            fsdp=fsdp_options,
            fsdp_config=fsdp_config,
            deepspeed=ds_config,
            ddp_find_unused_parameters=False,
            torch_compile=self.compile,
            max_grad_norm=self.max_grad_norm,
            **self.hf_training_args_overrides,
            **hf_training_args_overrides,
        )

    def _build_fsdp_config(self):  # src/train.py:126-136
        if self.fsdp_sharding == "no_shard":
            return "", None
        fsdp_options = [self.fsdp_sharding, "auto_wrap"]
        if self.fsdp_offload:
            fsdp_options += ["offload"]
        return fsdp_options, {"transformer_layer_cls_to_wrap": self.fsdp_layers_to_wrap}

    def _build_deepspeed_config(self) -> dict | None:  # src/train.py:138-215
        if self.zero_stage == "0":
            return None
        config: dict[str, Any] = {
            "fp16": {"enabled": "auto", "loss_scale": 0, "loss_scale_window": 1000, "initial_scale_power": 16,
                     "hysteresis": 2, "min_loss_scale": 1},
            "gradient_accumulation_steps": "auto",
            "gradient_clipping": "auto",
            "train_batch_size": "auto",
            "train_micro_batch_size_per_gpu": "auto",
        }
        if _is_adam(self.optimizer):
            config["optimizer"] = {"type": "Adam", "params": {"lr": "auto", "betas": "auto", "eps": "auto",
                                                             "weight_decay": "auto", "adam_w_mode": _is_adamw(self.optimizer)}}
        match self.zero_stage:
            case "1":
                config["zero_optimization"] = {"stage": 1}
            case "2":
                config["zero_optimization"] = {"stage": 2, "allgather_partitions": True, "allgather_bucket_size": 2e8,
                                               "overlap_comm": True, "reduce_scatter": True, "reduce_bucket_size": 2e8,
                                               "contiguous_gradients": True}
            case "3" | "3++":
                config["zero_optimization"] = {
                    "stage": 3, "overlap_comm": True, "contiguous_gradients": True, "sub_group_size": 1e9,
                    "reduce_bucket_size": "auto", "stage3_prefetch_bucket_size": "auto",
                    "stage3_param_persistence_threshold": "auto", "stage3_max_live_parameters": 1e9,
                    "stage3_max_reuse_distance": 1e9, "stage3_gather_16bit_weights_on_model_save": True}
                if self.zero_stage == "3++":
                    config["zero_optimization"].update(zero_quantized_weights=True,
                                                       zero_hpz_partition_size=torch.cuda.device_count(),
                                                       zero_quantized_gradients=True)
        if self.zero_offload_optimizer:
            config["zero_optimization"]["offload_optimizer"] = {"device": "cpu", "pin_memory": True}
        if self.zero_offload_params:
            config["zero_optimization"]["offload_param"] = {"device": "cpu", "pin_memory": True}
        return config
