"""ctypes binding of libb200pt.so (the C ABI declared in include/b200pt.h).

The library is the product: there is NO fallback. If it is missing, cannot be loaded, or the device is not sm_100,
every op raises. PyTorch is used only for device memory and streams.
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ["B200PT_LIB"]) if os.environ.get("B200PT_LIB") else _PKG / "libb200pt.so"  # override: kernel triage builds only
ABI_VERSION = 7

c_void_p, c_int, c_int64, c_float, c_size_t = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t


class GemmArgs(C.Structure):
    _fields_ = [
        ("M", c_int), ("N", c_int), ("K", c_int),
        ("A", c_void_p), ("lda", c_int64), ("a_mn", c_int),
        ("B", c_void_p), ("ldb", c_int64), ("b_mn", c_int),
        ("C", c_void_p), ("ldc", c_int64), ("c_fp32", c_int), ("accumulate", c_int),
        ("bias", c_void_p),
        ("residual", c_void_p), ("ldr", c_int64),
        ("gelu", c_int),
        ("alpha_dev", c_void_p),
        ("aux_out", c_void_p),
        ("dgelu_in", c_void_p),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("B", c_int), ("S", c_int), ("H", c_int), ("D", c_int),
        ("causal", c_int), ("scale", c_float),
        ("q", c_void_p), ("k", c_void_p), ("v", c_void_p),
        ("qkv_row_stride", c_int64), ("qkv_head_stride", c_int64),
        ("o", c_void_p), ("o_row_stride", c_int64), ("o_head_stride", c_int64),
        ("lse", c_void_p),
        ("d_o", c_void_p), ("delta", c_void_p),
        ("dq", c_void_p), ("dk", c_void_p), ("dv", c_void_p),
        ("dqkv_row_stride", c_int64), ("dqkv_head_stride", c_int64),
        ("p_scratch", c_void_p), ("ds_scratch", c_void_p),
        ("dropout_p", c_float), ("dropout_seed", C.c_uint64),
    ]


class AdamGroup(C.Structure):
    _fields_ = [
        ("lr", c_float), ("beta1", c_float), ("beta2", c_float), ("eps", c_float),
        ("weight_decay", c_float), ("bias_corr1", c_float), ("bias_corr2", c_float),
        ("adamw_mode", c_int),
    ]


# name -> (restype, argtypes); the single source of truth for "every symbol include/b200pt.h declares"
SIGNATURES = {
    "b200_abi_version": (c_int, []),
    "b200_last_error": (C.c_char_p, []),
    "b200_init": (c_int, [c_int]),
    "b200_layernorm_fwd": (c_int, [c_void_p] * 9 + [c_int, c_int, c_float, c_void_p]),
    "b200_layernorm_bwd_workspace_bytes": (c_size_t, [c_int, c_int]),
    "b200_layernorm_bwd": (c_int, [c_void_p] * 14 + [c_size_t, c_int, c_int, c_void_p]),
    "b200_gelu_fwd": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "b200_gelu_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "b200_rope_qk_inplace": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200_embedding_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "b200_embedding_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "b200_embedding_bwd_padding": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p]),
    "b200_embedding3_fwd": (c_int, [c_void_p] * 7 + [c_int, c_int, c_void_p]),
    "b200_roberta_position_ids": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p]),
    "b200_dropout": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_float, C.c_uint64, c_void_p]),
    "b200_count_valid": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    "b200_cross_entropy": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int64, c_int64, c_int, c_void_p]),
    "b200_mean_loss": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "b200_colsum_workspace_bytes": (c_size_t, [c_int]),
    "b200_colsum_bf16": (c_int, [c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "b200_gemm_bf16": (c_int, [C.POINTER(GemmArgs), c_void_p]),
    "b200_attention_fwd": (c_int, [C.POINTER(AttnArgs), c_void_p]),
    "b200_attention_bwd": (c_int, [C.POINTER(AttnArgs), c_void_p]),
    "b200_adam_step": (c_int, [c_void_p] * 5 + [c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int, C.POINTER(AdamGroup), c_int, c_void_p, c_int, c_void_p]),
    "b200_sumsq": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p]),
    "b200_clip_coef": (c_int, [c_void_p, c_float, c_void_p, c_void_p, c_void_p]),
    "b200_cast_f32_to_bf16": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "b200_scale_f32": (c_int, [c_void_p, c_size_t, c_void_p, c_float, c_void_p]),
}


class B200Error(RuntimeError):
    pass


_lib = None
_inited_devices: set[int] = set()


def load() -> C.CDLL:
    """dlopen libb200pt.so and bind every declared symbol. No compute, safe without a GPU."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise B200Error(
            f"{LIB_PATH} is missing: build it with `python -m multimodal_llm_pretraining_b200.csrc.build` "
            "(or __graft_entry__.build()). There is no fallback path."
        )
    lib = C.CDLL(str(LIB_PATH), mode=os.RTLD_NOW if hasattr(os, "RTLD_NOW") else 2)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    got = lib.b200_abi_version()
    if got != ABI_VERSION:
        raise B200Error(f"libb200pt ABI version {got} != expected {ABI_VERSION}: rebuild the library")
    _lib = lib
    return lib


def lib_for(device: torch.device | int) -> C.CDLL:
    """Library handle with b200_init done for `device`. Raises if no sm_100 GPU is present."""
    lib = load()
    idx = device if isinstance(device, int) else (device.index if device.index is not None else torch.cuda.current_device())
    if idx not in _inited_devices:
        if not torch.cuda.is_available():
            raise B200Error("libb200pt needs a CUDA device (sm_100a); none is available and there is no CPU fallback")
        with torch.cuda.device(idx):
            rc = lib.b200_init(idx)
        if rc != 0:
            raise B200Error(f"b200_init({idx}) failed: {lib.b200_last_error().decode()}")
        _inited_devices.add(idx)
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise B200Error(f"{what} failed ({rc}): {load().b200_last_error().decode()}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()
