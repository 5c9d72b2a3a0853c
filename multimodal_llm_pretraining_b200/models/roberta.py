"""RoBERTa model class — mirrors src/models/roberta.py:14-70."""
from __future__ import annotations

from typing import Any, Literal

import torch.optim
from torch import nn

from . import LanguageModelClass, RobertaT
from .configs import as_namespace, roberta_large_config_dict
from .pythia import SchedulerType


class RobertaModelClass(LanguageModelClass[RobertaT]):
    def config_dict(self) -> dict:
        return roberta_large_config_dict()

    def build_model(self, use_custom_kernels: bool = True) -> nn.Module:
        """src/models/roberta.py:15-18 (always eager attention there). True -> B200-native module."""
        cfg = self.config_dict()
        if use_custom_kernels:
            from ..modeling_roberta import B200RobertaForMaskedLM

            return B200RobertaForMaskedLM(as_namespace(cfg))
        from transformers import RobertaConfig, RobertaForMaskedLM

        return RobertaForMaskedLM(RobertaConfig(**cfg, attn_implementation="eager"))

    @property
    def batch_size(self) -> int:
        return 8192

    @property
    def training_steps(self) -> int:
        return 500000

    @property
    def mixed_precision(self) -> Literal[None, "bf16", "fp16"]:
        return "fp16"

    @property
    def optimizer(self) -> type[torch.optim.Optimizer]:  # src/models/roberta.py:32-34; swapped for the fused class by the trainer
        return torch.optim.Adam

    @property
    def optimizer_kwargs(self) -> dict[str, Any]:
        return {"lr": 4e-4, "betas": (0.9, 0.98), "weight_decay": 0.01}

    @property
    def scheduler_type(self):
        return SchedulerType("linear")

    @property
    def scheduler_kwargs(self) -> dict[str, Any]:
        return {"num_warmup_steps": 30_000}

    @property
    def max_grad_norm(self) -> float:
        return 0.0

    @property
    def hf_training_args(self) -> dict[str, Any]:
        return {}

    @property
    def fsdp_layers_to_wrap(self) -> list[str]:
        return ["RobertaLayer"]

    @property
    def vocab_size(self) -> int:
        return 50265

    @property
    def sequence_length(self) -> int:
        return 512
