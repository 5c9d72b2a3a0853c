"""Tensor-level wrappers over the C ABI (one function per entry point, no autograd, no fallbacks).

Every function enqueues on torch's current CUDA stream and returns immediately. Shapes/dtypes are validated here so the
C side only sees well-formed raw pointers.
"""

from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib
from ._lib import AdamGroup, AttnArgs, GemmArgs, check, ptr, stream_ptr

BF16 = torch.bfloat16
FP16 = torch.float16
HALF = (BF16, FP16)  # 16-bit element types: libb200pt.so / libb200pt_fp16.so (same entry points, see include/b200pt.h)
F32 = torch.float32

# number of libb200pt kernel launches issued through this module (bench.py reports it as `gpu_launches`)
LAUNCHES = 0
# optional GEMM profiler: a list that receives (flops, start_event, end_event) per GEMM call (bench.py roofline pass)
GEMM_PROFILE: list | None = None


def _count(n: int) -> None:
    global LAUNCHES
    LAUNCHES += n


def _L(t: torch.Tensor):
    """Library for the tensor's device and 16-bit element type (fp32-only ops run from the bf16 build)."""
    return _lib.lib_for(t.device, t.dtype if t.dtype in HALF else None)


def _req(cond: bool, msg: str) -> None:
    if not cond:
        raise ValueError(msg)


# ----------------------------------------------------------------------------------------------------- LayerNorm
def layernorm_fwd(x, gamma, beta, eps, gamma2=None, beta2=None):
    """x bf16 [rows, cols] -> (y, y2|None, mean, rstd)."""
    _req(x.dtype in HALF and x.is_contiguous() and x.dim() == 2, "layernorm_fwd: x must be contiguous bf16/fp16 [rows, cols]")
    _req(gamma.dtype == F32 and beta.dtype == F32, "layernorm_fwd: gamma/beta must be fp32")
    rows, cols = x.shape
    y = torch.empty_like(x)
    y2 = torch.empty_like(x) if gamma2 is not None else None
    mean = torch.empty(rows, dtype=F32, device=x.device)
    rstd = torch.empty(rows, dtype=F32, device=x.device)
    check(_L(x).b200_layernorm_fwd(ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(gamma2), ptr(beta2), ptr(y2), ptr(mean),
                                   ptr(rstd), rows, cols, float(eps), stream_ptr()), "b200_layernorm_fwd")
    _count(1)
    return y, y2, mean, rstd


_ws_cache: dict[tuple, torch.Tensor] = {}


def _workspace(device, nbytes: int) -> torch.Tensor:
    key = (device.index, torch.cuda.current_stream().cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def layernorm_bwd(x, mean, rstd, gamma, dy, dgamma, dbeta, gamma2=None, dy2=None, dgamma2=None, dbeta2=None, dres=None):
    """Returns dx (bf16). dgamma/dbeta (fp32) are accumulated in place."""
    rows, cols = x.shape
    _req(dy.dtype == x.dtype and x.dtype in HALF and dy.is_contiguous() and dy.shape == x.shape, "layernorm_bwd: dy must match x")
    _req(dgamma.dtype == F32 and dbeta.dtype == F32, "layernorm_bwd: dgamma/dbeta must be fp32")
    lib = _L(x)
    na = 2 if gamma2 is not None else 1
    nbytes = lib.b200_layernorm_bwd_workspace_bytes(cols, na)
    ws = _workspace(x.device, nbytes)
    dx = torch.empty_like(x)
    check(lib.b200_layernorm_bwd(ptr(x), ptr(mean), ptr(rstd), ptr(gamma), ptr(dy), ptr(gamma2), ptr(dy2), ptr(dres),
                                 ptr(dx), ptr(dgamma), ptr(dbeta), ptr(dgamma2), ptr(dbeta2), ptr(ws), ws.numel(), rows,
                                 cols, stream_ptr()), "b200_layernorm_bwd")
    _count(2)
    return dx


# ----------------------------------------------------------------------------------------------------- GELU / RoPE
def gelu_fwd(x):
    _req(x.dtype in HALF and x.is_contiguous(), "gelu_fwd: contiguous bf16/fp16 expected")
    y = torch.empty_like(x)
    check(_L(x).b200_gelu_fwd(ptr(x), ptr(y), x.numel(), stream_ptr()), "b200_gelu_fwd")
    _count(1)
    return y


def gelu_bwd(x, dy):
    _req(x.dtype in HALF and dy.dtype == x.dtype and x.is_contiguous() and dy.is_contiguous(), "gelu_bwd: contiguous bf16/fp16 expected")
    dx = torch.empty_like(x)
    check(_L(x).b200_gelu_bwd(ptr(x), ptr(dy), ptr(dx), x.numel(), stream_ptr()), "b200_gelu_bwd")
    _count(1)
    return dx


def rope_qk_inplace(qkv, cos, sin, B, S, nh, hd, rot, inverse=False):
    """qkv bf16 [B*S, nh*3*hd] packed per head as [q|k|v]; cos/sin fp32 [S, rot/2]."""
    _req(qkv.dtype in HALF and qkv.is_contiguous() and qkv.numel() == B * S * nh * 3 * hd, "rope: bad qkv")
    _req(cos.dtype == F32 and sin.dtype == F32 and cos.is_contiguous() and sin.is_contiguous(), "rope: cos/sin must be fp32")
    _req(cos.shape[0] >= S and cos.shape[1] == rot // 2, "rope: cos/sin must be [>=S, rot/2]")
    check(_L(qkv).b200_rope_qk_inplace(ptr(qkv), ptr(cos), ptr(sin), B, S, nh, hd, rot, int(inverse), stream_ptr()), "b200_rope_qk_inplace")
    _count(1)
    return qkv


# ----------------------------------------------------------------------------------------------------- Embedding
def embedding_fwd(ids, table):
    _req(ids.dtype == torch.int64 and ids.is_contiguous(), "embedding_fwd: ids must be contiguous int64")
    _req(table.dtype in HALF and table.is_contiguous(), "embedding_fwd: table must be contiguous bf16/fp16")
    T, h = ids.numel(), table.shape[1]
    out = torch.empty(T, h, dtype=table.dtype, device=table.device)
    check(_L(table).b200_embedding_fwd(ptr(ids), ptr(table), ptr(out), T, h, table.shape[0], stream_ptr()), "b200_embedding_fwd")
    _count(1)
    return out


def embedding3_fwd(ids0, table0, ids1=None, table1=None, ids2=None, table2=None):
    T, h = ids0.numel(), table0.shape[1]
    _req(table0.dtype in HALF and all(t is None or t.dtype == table0.dtype for t in (table1, table2)), "embedding3_fwd: tables must share a 16-bit dtype")
    out = torch.empty(T, h, dtype=table0.dtype, device=table0.device)
    check(_L(table0).b200_embedding3_fwd(ptr(ids0), ptr(table0), table0.shape[0], ptr(ids1), ptr(table1), ptr(ids2), ptr(table2), ptr(out), T, h, stream_ptr()), "b200_embedding3_fwd")
    _count(1)
    return out


def embedding_bwd(ids, dout, dtable, padding_idx=None):
    """dtable[ids[t]] += dout[t] (fp32 atomics); rows with ids == padding_idx are skipped (nn.Embedding padding_idx)."""
    _req(dout.dtype in HALF and dout.is_contiguous() and dtable.dtype == F32, "embedding_bwd: dout bf16/fp16, dtable fp32")
    T, h = ids.numel(), dtable.shape[1]
    if padding_idx is None:
        check(_L(dout).b200_embedding_bwd(ptr(ids), ptr(dout), ptr(dtable), T, h, dtable.shape[0], stream_ptr()), "b200_embedding_bwd")
    else:
        check(_L(dout).b200_embedding_bwd_padding(ptr(ids), ptr(dout), ptr(dtable), T, h, int(padding_idx), stream_ptr()), "b200_embedding_bwd_padding")
    _count(1)


def roberta_position_ids(ids, pad_id: int):
    """int64 [B,S] -> int64 [B,S]: cumsum(ids != pad) * (ids != pad) + pad (HF create_position_ids_from_input_ids)."""
    _req(ids.dtype == torch.int64 and ids.dim() == 2 and ids.is_contiguous(), "position_ids: ids must be contiguous int64 [B,S]")
    pos = torch.empty_like(ids)
    check(_L(ids).b200_roberta_position_ids(ptr(ids), ptr(pos), ids.shape[0], ids.shape[1], int(pad_id), stream_ptr()), "b200_roberta_position_ids")
    _count(1)
    return pos


def dropout(x, p: float, seed: int, residual=None, out=None):
    """out = dropout(x) (+ residual), mask = f(seed, element index); call again on the gradient with the same seed for backward."""
    _req(x.dtype in HALF and x.is_contiguous() and x.numel() % 8 == 0, "dropout: contiguous bf16/fp16 with numel % 8 == 0")
    _req(residual is None or (residual.dtype == x.dtype and residual.is_contiguous() and residual.numel() == x.numel()), "dropout: bad residual")
    out = torch.empty_like(x) if out is None else out
    check(_L(x).b200_dropout(ptr(x), ptr(residual), ptr(out), x.numel(), float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, stream_ptr()), "b200_dropout")
    _count(1)
    return out


# ----------------------------------------------------------------------------------------------------- Cross entropy
def cross_entropy_(logits, labels, V=None, ignore_index=-100, write_grad=True, grad_scale=None):
    """logits bf16/fp16 [T, ld] (overwritten by dlogits/n_valid when write_grad). Returns (loss scalar fp32 tensor, n_valid int tensor).
    grad_scale (device fp32 scalar, optional): multiplied into dlogits — the fp16 loss scale. Labels outside [0, V) that are
    not ignore_index trap the kernel (torch raises a device-side assert for them)."""
    _req(logits.dtype in HALF and logits.dim() == 2 and logits.stride(1) == 1, "cross_entropy: logits must be bf16/fp16 [T, ld]")
    _req(grad_scale is None or (grad_scale.dtype == F32 and grad_scale.numel() == 1), "cross_entropy: grad_scale must be a device fp32 scalar")
    _req(labels.dtype == torch.int64 and labels.is_contiguous(), "cross_entropy: labels must be contiguous int64")
    T, ld = logits.shape[0], logits.stride(0)
    V = logits.shape[1] if V is None else V
    lib = _L(logits)
    dev = logits.device
    n_valid = torch.empty(1, dtype=torch.int32, device=dev)
    row_loss = torch.empty(T, dtype=F32, device=dev)
    loss = torch.empty((), dtype=F32, device=dev)
    s = stream_ptr()
    check(lib.b200_count_valid(ptr(labels), T, ignore_index, V, ptr(n_valid), s), "b200_count_valid")
    _count(1)
    check(lib.b200_cross_entropy(ptr(logits), ptr(labels), ptr(row_loss), ptr(n_valid), T, V, ld, ignore_index, int(write_grad), ptr(grad_scale), s), "b200_cross_entropy")
    _count(1)
    check(lib.b200_mean_loss(ptr(row_loss), ptr(n_valid), T, ptr(loss), s), "b200_mean_loss")
    _count(1)
    return loss, n_valid


def colsum_(x, out, out2=None, scale=None):
    """out[c] (fp32) += s * sum_r x[r, c] for bf16 x [rows, cols]; out2 (optional) gets the same increment; scale = device fp32 scalar s."""
    _req(out2 is None or (out2.dtype == F32 and out2.numel() == out.numel()), "colsum: bad out2")
    _req(scale is None or (scale.dtype == F32 and scale.numel() == 1), "colsum: scale must be a device fp32 scalar")
    _req(x.dtype in HALF and x.dim() == 2 and x.stride(1) == 1 and out.dtype == F32 and out.numel() == x.shape[1], "colsum: bad tensors")
    lib = _L(x)
    rows, cols = x.shape
    ws = _workspace(x.device, lib.b200_colsum_workspace_bytes(cols))
    check(lib.b200_colsum_bf16(ptr(x), rows, cols, x.stride(0), ptr(out), ptr(out2), ptr(scale), ptr(ws), ws.numel(), stream_ptr()), "b200_colsum_bf16")
    _count(2)
    return out


# ----------------------------------------------------------------------------------------------------- GEMM
def gemm(A, B, *, a_mn=False, b_mn=False, out=None, out_dtype=None, accumulate=False, bias=None, residual=None,
         gelu=False, alpha=None, aux_out=None, dgelu_in=None, dropout_p: float = 0.0, dropout_seed: int = 0, colsum_out=None):
    """C[M,N] = epi(alpha * A·Bᵀ).  A is [M,K] (a_mn=False) or [K,M] (a_mn=True); B is [N,K] or [K,N] (b_mn=True).
    dropout_p > 0 (needs residual): C = dropout(alpha * A·Bᵀ + bias) + residual with K.dropout's mask of (dropout_seed, element)."""
    _req(A.dtype in HALF and B.dtype == A.dtype and A.dim() == 2 and B.dim() == 2, "gemm: A,B must be 2-D bf16 (or both fp16)")
    E = A.dtype
    out_dtype = E if out_dtype is None else out_dtype
    _req(A.stride(1) == 1 and B.stride(1) == 1, "gemm: A,B must have unit inner stride")
    if a_mn:
        K, M = A.shape
    else:
        M, K = A.shape
    if b_mn:
        Kb, N = B.shape
    else:
        N, Kb = B.shape
    _req(K == Kb, f"gemm: reduction dims differ ({K} vs {Kb})")
    if out is None:
        _req(not accumulate, "gemm: accumulate needs out")
        out = torch.empty(M, N, dtype=out_dtype, device=A.device)
    _req(out.shape == (M, N) and out.stride(1) == 1 and out.dtype in (E, F32), "gemm: bad out")
    a = GemmArgs()
    a.M, a.N, a.K = M, N, K
    a.A, a.lda, a.a_mn = ptr(A), A.stride(0), int(a_mn)
    a.B, a.ldb, a.b_mn = ptr(B), B.stride(0), int(b_mn)
    a.C, a.ldc, a.c_fp32, a.accumulate = ptr(out), out.stride(0), int(out.dtype == F32), int(accumulate)
    if bias is not None:
        _req(bias.dtype == F32 and bias.numel() == N and bias.is_contiguous(), "gemm: bias must be fp32 [N]")
        a.bias = ptr(bias)
    ldr = 0
    if residual is not None:
        _req(residual.dtype == E and residual.shape == (M, N) and residual.stride(1) == 1, "gemm: bad residual")
        a.residual, ldr = ptr(residual), residual.stride(0)
    if dgelu_in is not None:
        _req(dgelu_in.dtype == E and dgelu_in.shape == (M, N) and dgelu_in.stride(1) == 1, "gemm: bad dgelu_in")
        _req(residual is None or residual.stride(0) == dgelu_in.stride(0), "gemm: residual and dgelu_in must share a pitch")
        a.dgelu_in, ldr = ptr(dgelu_in), dgelu_in.stride(0)
    a.ldr = ldr
    a.gelu = int(gelu)
    if alpha is not None:
        _req(alpha.dtype == F32 and alpha.numel() == 1, "gemm: alpha must be a device fp32 scalar")
        a.alpha_dev = ptr(alpha)
    if aux_out is not None:
        _req(aux_out.dtype == E and aux_out.shape == (M, N) and aux_out.stride(0) == out.stride(0) and out.dtype == E, "gemm: aux_out must match a 16-bit out")
        a.aux_out = ptr(aux_out)
    if colsum_out is not None:
        _req(dgelu_in is not None and colsum_out.dtype == F32 and colsum_out.numel() == N and colsum_out.is_contiguous(),
             "gemm: colsum_out (fp32 [N], accumulated) rides in the dGELU dgrad epilogue only")
        a.colsum_out = ptr(colsum_out)
    if dropout_p > 0.0:
        _req(residual is not None and out.is_contiguous(), "gemm: fused dropout needs a residual and a contiguous output")
        a.dropout_p, a.dropout_seed = float(dropout_p), int(dropout_seed) & 0xFFFFFFFFFFFFFFFF
    prof = GEMM_PROFILE
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(_L(A).b200_gemm_bf16(C.byref(a), stream_ptr()), "b200_gemm_bf16")
    _count(1)
    if prof is not None:
        e1.record()
        prof.append((2.0 * M * N * K, e0, e1))
    return out


# ----------------------------------------------------------------------------------------------------- Attention
def _attn_args(q, k, v, o, lse, B, S, H, D, causal, scale):
    a = AttnArgs()
    a.B, a.S, a.H, a.D = B, S, H, D
    a.causal = int(causal)
    a.scale = float(scale)
    for t in (q, k, v):
        _req(t.dtype in HALF and t.dtype == q.dtype and t.dim() == 4 and t.shape == (B, S, H, D) and t.stride(3) == 1, "attention: q,k,v must be bf16/fp16 views [B,S,H,D] with unit inner stride")
        _req(t.stride(0) == S * t.stride(1), "attention: batch stride must equal S * token stride")
        _req(t.stride(1) == q.stride(1) and t.stride(2) == q.stride(2), "attention: q,k,v must share strides")
    a.q, a.k, a.v = ptr(q), ptr(k), ptr(v)
    a.qkv_row_stride, a.qkv_head_stride = q.stride(1), q.stride(2)
    _req(o.dtype == q.dtype and o.shape == (B, S, H, D) and o.stride(3) == 1 and o.stride(0) == S * o.stride(1), "attention: bad o")
    a.o, a.o_row_stride, a.o_head_stride = ptr(o), o.stride(1), o.stride(2)
    _req(lse.dtype == F32 and lse.shape == (B, H, S) and lse.is_contiguous(), "attention: lse must be fp32 [B,H,S]")
    a.lse = ptr(lse)
    return a


USE_SCORE_SCRATCH = True
_score_cache: dict[tuple, tuple[torch.Tensor, torch.Tensor]] = {}


def _score_scratch(device, Z: int, S: int, dtype=BF16):
    """Persistent bf16 [Z, S, S] P and dS scratch for the head_dim-256 backward (reused by every layer and step)."""
    key = (device.index, torch.cuda.current_stream().cuda_stream, Z, S, dtype)
    if key not in _score_cache:
        _score_cache.clear()  # one shape at a time: these are GB-sized
        _score_cache[key] = (torch.empty(Z, S, S, dtype=dtype, device=device), torch.empty(Z, S, S, dtype=dtype, device=device))
    return _score_cache[key]


def attention_fwd(q, k, v, causal, scale=None, dropout_p: float = 0.0, dropout_seed: int = 0):
    """q,k,v: bf16 views [B,S,H,D] (any token/head strides, e.g. slices of a packed qkv buffer). Returns (o [B,S,H,D], lse [B,H,S]).
    dropout_p > 0 drops softmax probabilities with the counter-based mask of (dropout_seed, head, query, key)."""
    B, S, H, D = q.shape
    scale = D ** -0.5 if scale is None else scale
    o = torch.empty(B, S, H, D, dtype=q.dtype, device=q.device)
    lse = torch.empty(B, H, S, dtype=F32, device=q.device)
    a = _attn_args(q, k, v, o, lse, B, S, H, D, causal, scale)
    a.dropout_p, a.dropout_seed = float(dropout_p), int(dropout_seed) & 0xFFFFFFFFFFFFFFFF
    check(_L(q).b200_attention_fwd(C.byref(a), stream_ptr()), "b200_attention_fwd")
    _count(1)
    return o, lse


def attention_bwd(q, k, v, o, lse, d_o, dq, dk, dv, causal, scale=None, dropout_p: float = 0.0, dropout_seed: int = 0):
    """dq,dk,dv: bf16 views [B,S,H,D] sharing strides (e.g. slices of a packed dqkv buffer); written, not accumulated.
    (dropout_p, dropout_seed) must be the forward's."""
    B, S, H, D = q.shape
    scale = D ** -0.5 if scale is None else scale
    a = _attn_args(q, k, v, o, lse, B, S, H, D, causal, scale)
    a.dropout_p, a.dropout_seed = float(dropout_p), int(dropout_seed) & 0xFFFFFFFFFFFFFFFF
    _req(d_o.dtype == q.dtype and d_o.shape == o.shape and d_o.stride() == o.stride(), "attention_bwd: dO must match O")
    for t in (dq, dk, dv):
        _req(t.dtype == q.dtype and t.shape == (B, S, H, D) and t.stride(3) == 1 and t.stride(0) == S * t.stride(1), "attention_bwd: bad dq/dk/dv")
        _req(t.stride(1) == dq.stride(1) and t.stride(2) == dq.stride(2), "attention_bwd: dq,dk,dv must share strides")
    delta = torch.empty(B, H, S, dtype=F32, device=q.device)
    a.d_o, a.delta = ptr(d_o), ptr(delta)
    n_launch = 3
    if D == 256 and S % 256 == 0 and USE_SCORE_SCRATCH and dropout_p == 0.0:
        # head_dim 256: P / dS tiles go through HBM scratch and dK / dV become batched GEMMs (b200pt.h, b200_attn_args)
        ps, dss = _score_scratch(q.device, B * H, S, q.dtype)
        a.p_scratch, a.ds_scratch = ptr(ps), ptr(dss)
        n_launch = 4
    a.dq, a.dk, a.dv = ptr(dq), ptr(dk), ptr(dv)
    a.dqkv_row_stride, a.dqkv_head_stride = dq.stride(1), dq.stride(2)
    check(_L(q).b200_attention_bwd(C.byref(a), stream_ptr()), "b200_attention_bwd")
    _count(n_launch)
    return dq, dk, dv


# ----------------------------------------------------------------------------------------------------- Optimizer
def adam_step(p, g, m, v, p_bf16, state_base, chunk_start, chunk_len, chunk_group, groups, grad_scale=None, zero_grad=False,
              chunk_state=None, skip_flag=None, g_packed=False, p_packed=False):
    """p_bf16: the 16-bit compute copy (bf16 or fp16; its dtype selects the library build). skip_flag (device int32, optional):
    non-zero leaves p, m, v untouched (fp16 overflow step). g_packed / p_packed: g / p are packed shard buffers indexed like m / v (ZeRO-2 gradients / a sharded fp32 master)."""
    n_chunks = chunk_start.numel()
    arr = (AdamGroup * len(groups))()
    for i, gdict in enumerate(groups):
        arr[i].lr, arr[i].beta1, arr[i].beta2, arr[i].eps = gdict["lr"], gdict["beta1"], gdict["beta2"], gdict["eps"]
        arr[i].weight_decay, arr[i].bias_corr1, arr[i].bias_corr2 = gdict["weight_decay"], gdict["bias_corr1"], gdict["bias_corr2"]
        arr[i].adamw_mode = int(gdict["adamw_mode"])
    lib = _L(p_bf16) if p_bf16 is not None else _L(p)
    check(lib.b200_adam_step(ptr(p), ptr(g), ptr(m), ptr(v), ptr(p_bf16), int(state_base), ptr(chunk_start), ptr(chunk_len),
                             ptr(chunk_group), ptr(chunk_state), n_chunks, arr, len(groups), ptr(grad_scale), int(zero_grad),
                             ptr(skip_flag), int(g_packed), int(p_packed), stream_ptr()), "b200_adam_step")
    _count(1)


def sumsq_(x, out):
    """out (fp32 scalar tensor) += sum(x^2); deterministic (fixed-order two-pass reduction, no atomics)."""
    _req(x.dtype == F32 and x.is_contiguous() and out.dtype == F32, "sumsq: fp32 expected")
    lib = _L(x)
    ws = _workspace(x.device, lib.b200_sumsq_workspace_bytes())
    check(lib.b200_sumsq(ptr(x), x.numel(), ptr(out), ptr(ws), ws.numel(), stream_ptr()), "b200_sumsq")
    _count(2)
    return out


def sumsq_chunks_(x, chunk_start, chunk_len, out, partials=None):
    """out += sum of squares of x over the chunks (chunk_start int64, chunk_len int32 device arrays): the slices a ZeRO rank owns."""
    _req(x.dtype == F32 and x.is_contiguous() and out.dtype == F32, "sumsq_chunks: fp32 expected")
    n = chunk_start.numel()
    if partials is None or partials.numel() < n:
        partials = torch.empty(max(n, 1), dtype=F32, device=x.device)
    check(_L(x).b200_sumsq_chunks(ptr(x), ptr(chunk_start), ptr(chunk_len), n, ptr(out), ptr(partials), stream_ptr()), "b200_sumsq_chunks")
    _count(2)
    return out


def clip_coef(sumsq, max_norm, loss_scale=None, found_inf=None):
    """(norm, coef): norm = sqrt(sumsq) / loss_scale, coef = min(1, max_norm / (norm + 1e-6)) / loss_scale. found_inf (device
    int32, optional) receives 1 when sumsq is inf / nan."""
    norm = torch.empty((), dtype=F32, device=sumsq.device)
    coef = torch.empty((), dtype=F32, device=sumsq.device)
    check(_L(sumsq).b200_clip_coef(ptr(sumsq), float(max_norm if max_norm is not None else 0.0), ptr(loss_scale), ptr(norm), ptr(coef),
                                   ptr(found_inf), stream_ptr()), "b200_clip_coef")
    _count(1)
    return norm, coef


def loss_scale_update(scale, growth_tracker, hysteresis_left, found_inf, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000,
                      min_scale=1.0, hysteresis=1):
    """Device-side dynamic loss-scale update (b200pt.h: b200_loss_scale_update)."""
    _req(scale.dtype == F32 and growth_tracker.dtype == torch.int32 and hysteresis_left.dtype == torch.int32 and found_inf.dtype == torch.int32,
         "loss_scale_update: scale fp32, counters / flag int32")
    check(_L(scale).b200_loss_scale_update(ptr(scale), ptr(growth_tracker), ptr(hysteresis_left), ptr(found_inf), float(growth_factor),
                                           float(backoff_factor), int(growth_interval), float(min_scale), int(hysteresis), stream_ptr()),
          "b200_loss_scale_update")
    _count(1)


def cast_f32_to_bf16(src, dst):
    """dst (bf16 or fp16, selects the library build) = cast(src fp32)."""
    _req(src.dtype == F32 and dst.dtype in HALF and src.numel() == dst.numel() and src.is_contiguous() and dst.is_contiguous(), "cast: bad tensors")
    check(_L(dst).b200_cast_f32_to_bf16(ptr(src), ptr(dst), src.numel(), stream_ptr()), "b200_cast_f32_to_bf16")
    _count(1)
    return dst


def scale_f32_(x, scale_dev=None, scale_host=1.0):
    _req(x.dtype == F32 and x.is_contiguous(), "scale_f32: contiguous fp32 expected")
    check(_L(x).b200_scale_f32(ptr(x), x.numel(), ptr(scale_dev), float(scale_host), stream_ptr()), "b200_scale_f32")
    _count(1)
    return x
