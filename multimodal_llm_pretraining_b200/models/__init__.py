"""Model registry — mirrors src/models/__init__.py of the reference for the in-scope families (Pythia, RoBERTa).

Same surface: `ModelT`, `BaseModelClass` (build_model / batch_size / training_steps / mixed_precision / optimizer /
optimizer_kwargs / scheduler_type / scheduler_kwargs / max_grad_norm / hf_training_args / fsdp_layers_to_wrap /
load_dummy_dataset), `LanguageModelClass`, `get_model_class` (src/models/__init__.py:67-181,240-296).
`build_model(use_custom_kernels=True)` returns the B200-native nn.Module; `False` returns the stock HuggingFace eager
module (the reference's "naive" path, used as the live oracle in tests).  Out-of-scope families (Mamba, ViT, ConvNeXt,
LLaVA, ViLT) raise NotImplementedError.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Generic, Literal, TypeVar, get_args

import torch.optim
from torch import nn
from torch.utils.data import Dataset

from ..benchmarking.data import DummyTextModelingDataset

RobertaT = Literal["roberta"]
PythiaT = Literal[
    "pythia-14m", "pythia-31m", "pythia-70m", "pythia-160m", "pythia-410m", "pythia-1b", "pythia-1.4b",
    "pythia-2.8b", "pythia-6.9b", "pythia-12b",
]
ModelT = Literal[RobertaT, PythiaT]

T = TypeVar("T")


class BaseModelClass(ABC, Generic[T]):
    """Define models and hyper-parameters using this class (src/models/__init__.py:67-162)."""

    def __init__(self, model_type: T) -> None:
        self.model_type: T = model_type

    @abstractmethod
    def build_model(self, use_custom_kernels: bool = True) -> nn.Module:
        raise NotImplementedError

    @property
    def supports_activation_checkpointing(self) -> bool:
        return True

    @property
    def supports_compilation(self) -> bool:
        return True

    @property
    @abstractmethod
    def batch_size(self) -> int: ...

    @property
    @abstractmethod
    def training_steps(self) -> int: ...

    @property
    @abstractmethod
    def mixed_precision(self) -> Literal[None, "bf16", "fp16"]: ...

    @property
    @abstractmethod
    def optimizer(self) -> type[torch.optim.Optimizer]: ...

    @property
    @abstractmethod
    def optimizer_kwargs(self) -> dict[str, Any]: ...

    @property
    @abstractmethod
    def scheduler_type(self): ...

    @property
    @abstractmethod
    def scheduler_kwargs(self) -> dict[str, Any]: ...

    @property
    @abstractmethod
    def max_grad_norm(self) -> float: ...

    @property
    @abstractmethod
    def hf_training_args(self) -> dict[str, Any]: ...

    @property
    @abstractmethod
    def fsdp_layers_to_wrap(self) -> list[str]: ...

    @abstractmethod
    def load_dummy_dataset(self) -> Dataset: ...


class LanguageModelClass(Generic[T], BaseModelClass[T]):
    """src/models/__init__.py:165-181."""

    @property
    @abstractmethod
    def vocab_size(self) -> int: ...

    @property
    @abstractmethod
    def sequence_length(self) -> int: ...

    def load_dummy_dataset(self, num_samples: int = 50_000, seed: int | None = None) -> Dataset:
        return DummyTextModelingDataset(vocab_size=self.vocab_size, sequence_length=self.sequence_length,
                                        num_samples=num_samples, seed=seed)


def get_model_class(model_type: str) -> BaseModelClass:
    """src/models/__init__.py:240-296 (dispatch on the ModelT literal)."""
    if model_type in get_args(PythiaT):
        from .pythia import PythiaModelClass

        return PythiaModelClass(model_type)
    if model_type in get_args(RobertaT):
        from .roberta import RobertaModelClass

        return RobertaModelClass(model_type)
    raise NotImplementedError(
        f"model {model_type!r} is outside this build's hot-path scope (Pythia / RoBERTa pretraining step; SURVEY.md §8)")
