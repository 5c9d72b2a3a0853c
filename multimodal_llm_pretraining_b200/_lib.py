"""ctypes binding of libb200pt.so / libb200pt_fp16.so (the C ABI declared in include/b200pt.h; one build per 16-bit
element type, same entry points).

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
LIB_PATH_FP16 = _PKG / "libb200pt_fp16.so"
ABI_VERSION = 8

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
        ("dropout_p", c_float), ("dropout_seed", C.c_uint64),
        ("colsum_out", c_void_p),
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
    "b200_elem_dtype": (c_int, []),
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
    "b200_embedding3_fwd": (c_int, [c_void_p, c_void_p, c_int] + [c_void_p] * 5 + [c_int, c_int, c_void_p]),
    "b200_roberta_position_ids": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p]),
    "b200_dropout": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_float, C.c_uint64, c_void_p]),
    "b200_count_valid": (c_int, [c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p]),
    "b200_cross_entropy": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "b200_mean_loss": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "b200_colsum_workspace_bytes": (c_size_t, [c_int]),
    "b200_colsum_bf16": (c_int, [c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "b200_gemm_bf16": (c_int, [C.POINTER(GemmArgs), c_void_p]),
    "b200_attention_fwd": (c_int, [C.POINTER(AttnArgs), c_void_p]),
    "b200_attention_bwd": (c_int, [C.POINTER(AttnArgs), c_void_p]),
    "b200_adam_step": (c_int, [c_void_p] * 5 + [c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int, C.POINTER(AdamGroup), c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p]),
    "b200_sumsq_workspace_bytes": (c_size_t, []),
    "b200_sumsq": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p, c_size_t, c_void_p]),
    "b200_sumsq_chunks": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "b200_clip_coef": (c_int, [c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200_loss_scale_update": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_int, c_float, c_int, c_void_p]),
    "b200_cast_f32_to_bf16": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "b200_scale_f32": (c_int, [c_void_p, c_size_t, c_void_p, c_float, c_void_p]),
}


class B200Error(RuntimeError):
    pass


_libs: dict[str, C.CDLL] = {}
_inited_devices: set[tuple[str, int]] = set()


def _variant(dtype) -> str:
    if dtype is None or dtype == torch.bfloat16 or dtype == "bf16":
        return "bf16"
    if dtype == torch.float16 or dtype == "fp16":
        return "fp16"
    raise B200Error(f"libb200pt is built for bf16 and fp16 tensors, not {dtype}")


def load(dtype=None) -> C.CDLL:
    """dlopen the library for `dtype` (bf16 default, fp16) and bind every declared symbol. No compute, safe without a GPU."""
    var = _variant(dtype)
    if var in _libs:
        return _libs[var]
    path = LIB_PATH if var == "bf16" else LIB_PATH_FP16
    if not path.exists():
        raise B200Error(
            f"{path} is missing: build it with `python -m multimodal_llm_pretraining_b200.csrc.build` "
            "(or __graft_entry__.build()). There is no fallback path."
        )
    lib = C.CDLL(str(path), mode=os.RTLD_NOW if hasattr(os, "RTLD_NOW") else 2)  # RTLD_LOCAL: the two builds share symbol names
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    got = lib.b200_abi_version()
    if got != ABI_VERSION:
        raise B200Error(f"{path.name} ABI version {got} != expected {ABI_VERSION}: rebuild the library")
    if lib.b200_elem_dtype() != (0 if var == "bf16" else 1):
        raise B200Error(f"{path.name} reports element type {lib.b200_elem_dtype()}, expected the {var} build")
    _libs[var] = lib
    return lib


def lib_for(device: torch.device | int, dtype=None) -> C.CDLL:
    """Library handle (for `dtype`) with b200_init done for `device`. Raises if no sm_100 GPU is present."""
    var = _variant(dtype)
    lib = load(var)
    idx = device if isinstance(device, int) else (device.index if device.index is not None else torch.cuda.current_device())
    if (var, idx) not in _inited_devices:
        if not torch.cuda.is_available():
            raise B200Error("libb200pt needs a CUDA device (sm_100a); none is available and there is no CPU fallback")
        with torch.cuda.device(idx):
            rc = lib.b200_init(idx)
        if rc != 0:
            raise B200Error(f"b200_init({idx}) failed: {lib.b200_last_error().decode()}")
        _inited_devices.add((var, idx))
    return lib


def check(rc: int, what: str, lib: C.CDLL | None = None) -> None:
    if rc != 0:
        msgs = [l.b200_last_error().decode() for l in ([lib] if lib is not None else list(_libs.values()))]
        raise B200Error(f"{what} failed ({rc}): {' | '.join(m for m in msgs if m)}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()
