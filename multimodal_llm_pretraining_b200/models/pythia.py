"""Pythia model class — mirrors src/models/pythia.py:14-98 (hyper-parameters verbatim in meaning; cited per property)."""
from __future__ import annotations

from typing import Any, Literal

import torch.optim
from torch import nn

from . import LanguageModelClass, PythiaT
from .configs import as_namespace, pythia_config_dict


class SchedulerType(str):
    """Stand-in for transformers.SchedulerType: a str with `.value` (src/train.py:108 uses `.value`)."""

    @property
    def value(self) -> str:
        return str(self)


class PythiaModelClass(LanguageModelClass[PythiaT]):
    def config_dict(self) -> dict:
        return pythia_config_dict(self.model_type)

    def build_model(self, use_custom_kernels: bool = True) -> nn.Module:
        """src/models/pythia.py:15-22. use_custom_kernels=True -> B200-native module (in place of HF + sdpa);
        False -> stock HF GPTNeoXForCausalLM with eager attention (the reference's naive path)."""
        cfg = self.config_dict()
        if use_custom_kernels:
            from ..modeling_gpt_neox import B200GPTNeoXForCausalLM

            return B200GPTNeoXForCausalLM(as_namespace(cfg))
        from transformers import GPTNeoXConfig, GPTNeoXForCausalLM

        return GPTNeoXForCausalLM(GPTNeoXConfig(**cfg, attn_implementation="eager"))

    @property
    def batch_size(self) -> int:  # :24-26
        return 1024

    @property
    def training_steps(self) -> int:  # :28-30
        return 143000

    @property
    def mixed_precision(self) -> Literal[None, "bf16", "fp16"]:  # :32-41
        return "bf16" if self.model_type == "pythia-1b" else "fp16"

    @property
    def optimizer(self) -> type[torch.optim.Optimizer]:
        """src/models/pythia.py:43-45: torch.optim.Adam (L2-coupled). For a B200 module the trainer swaps it for the fused
        equivalent (optim.fused_optimizer_class); the naive HF module keeps the torch class, as in the reference."""
        return torch.optim.Adam

    @property
    def optimizer_kwargs(self) -> dict[str, Any]:  # :47-67
        match self.model_type:
            case "pythia-14m" | "pythia-31m" | "pythia-70m":
                lr = 1.0e-3
            case "pythia-160m":
                lr = 6.0e-4
            case "pythia-410m" | "pythia-1b":
                lr = 3.0e-4
            case "pythia-1.4b":
                lr = 2.0e-4
            case "pythia-2.8b":
                lr = 1.6e-4
            case "pythia-6.9b" | "pythia-12b":
                lr = 1.2e-4
        return {"lr": lr, "betas": (0.9, 0.95), "eps": 1e-8, "weight_decay": 0.01}

    @property
    def scheduler_type(self):  # :69-71
        return SchedulerType("cosine_with_min_lr")

    @property
    def scheduler_kwargs(self) -> dict[str, Any]:  # :73-78
        return {"num_warmup_steps": int(0.01 * self.training_steps), "min_lr_rate": 0.1}

    @property
    def max_grad_norm(self) -> float:  # :80-82
        return 1.0

    @property
    def hf_training_args(self) -> dict[str, Any]:
        return {}

    @property
    def fsdp_layers_to_wrap(self) -> list[str]:
        return ["GPTNeoXLayer"]

    @property
    def vocab_size(self) -> int:  # :92-94 (token range of the dummy data)
        return 50304

    @property
    def sequence_length(self) -> int:  # :96-98 — 2049 tokens in, 2048 predicted
        return 2049
