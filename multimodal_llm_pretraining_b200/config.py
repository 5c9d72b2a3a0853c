"""BaseConfig / TrainingConfig — mirrors experiments/config.py:12-101 (the CLI-facing constructor of TrainingClass)."""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass
from typing import Literal

from .gpus import GpuT, ampere_or_newer_gpu
from .models import BaseModelClass, ModelT, get_model_class
from .train import TrainingClass

ShardingT = Literal["", "fsdp_shard_grad_op", "fsdp_full_shard", "fsdp_hybrid_shard_zero2", "fsdp_hybrid_shard",
                    "zero_1", "zero_2", "zero_3", "zero_3++"]


@dataclass
class BaseConfig:
    num_nodes: int
    gpus_per_node: int
    gpu_type: GpuT
    model: ModelT

    def ampere_or_newer_gpu(self) -> bool:
        return ampere_or_newer_gpu(self.gpu_type)

    def model_class(self) -> BaseModelClass:
        return get_model_class(model_type=self.model)

    def __str__(self) -> str:  # deterministic, like the reference's TangoStringHash (experiments/utils/__tango__.py:34-37)
        return f"{type(self).__name__}({', '.join(f'{f.name}={getattr(self, f.name)!r}' for f in dataclasses.fields(self))})"


@dataclass
class TrainingConfig(BaseConfig):
    free_lunch: bool = False
    activation_checkpointing: bool = False
    sharding: ShardingT = ""
    offloading: bool = False

    def training_class(self, **training_class_overrides) -> TrainingClass:
        model_class = self.model_class()
        if self.free_lunch:  # experiments/config.py:43-48
            tf32 = self.ampere_or_newer_gpu()
            compile = model_class.supports_compilation
        else:
            tf32, compile = False, False
        fsdp_sharding, fsdp_layers_to_wrap, fsdp_offload = "no_shard", [], False
        zero_stage, zero_offload_optimizer, zero_offload_params = "0", False, False
        if self.sharding.startswith("fsdp_"):
            fsdp_sharding = self.sharding[len("fsdp_"):]
            fsdp_layers_to_wrap = model_class.fsdp_layers_to_wrap
            fsdp_offload = bool(self.offloading)
        elif self.sharding.startswith("zero_"):
            zero_stage = self.sharding[len("zero_"):]
            if self.offloading:
                zero_offload_optimizer = True
                zero_offload_params = zero_stage in ["3", "3++"]
        tc = TrainingClass(
            num_training_steps=model_class.training_steps, micro_batch_size=1, gradient_accumulation_steps=1,
            gradient_checkpointing=self.activation_checkpointing,
            bf16=(model_class.mixed_precision == "bf16"), fp16=(model_class.mixed_precision == "fp16"),
            tf32=tf32, compile=compile, optimizer=model_class.optimizer, optimizer_kwargs=model_class.optimizer_kwargs,
            scheduler_type=model_class.scheduler_type, scheduler_kwargs=model_class.scheduler_kwargs,
            fsdp_sharding=fsdp_sharding, fsdp_layers_to_wrap=fsdp_layers_to_wrap, fsdp_offload=fsdp_offload,
            zero_stage=zero_stage, zero_offload_optimizer=zero_offload_optimizer, zero_offload_params=zero_offload_params,
            max_grad_norm=model_class.max_grad_norm, hf_training_args_overrides=model_class.hf_training_args,
        )
        return dataclasses.replace(tc, **training_class_overrides)
