"""TEST INFRASTRUCTURE ONLY: torch statements (fp32 math, 16-bit rounding at the same places) of the tensor-level contracts in
multimodal_llm_pretraining_b200/kernels.py, one per wrapper, so that the HOST logic above the C ABI — the hand-scheduled
forward / backward of the real modules, gradient / parameter routing, bucket hooks, the fused-optimizer chunk tables, the
TrainEngine strategies — can be executed on CPU tensors (and over gloo with world_size 2) in `-m "not gpu"` tests.

Nothing in the product imports this file; `install()` monkey-patches the `kernels` module of the CURRENT test process. The kernels
themselves are checked against references on the GPU (tests/test_kernels_gpu.py, test_attention_gpu.py); what this file buys is
coverage of everything around them where no GPU is present.
"""
from __future__ import annotations

import math

import torch

F32 = torch.float32


def _gelu(x):
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def _dgelu(x):
    return 0.5 * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0)))) + x * torch.exp(-0.5 * x * x) * (1.0 / math.sqrt(2.0 * math.pi))


# ----------------------------------------------------------------------------------------------------- LayerNorm
def layernorm_fwd(x, gamma, beta, eps, gamma2=None, beta2=None):
    xf = x.float()
    mean = xf.mean(1)
    var = xf.var(1, unbiased=False)
    rstd = torch.rsqrt(var + eps)
    xhat = (xf - mean[:, None]) * rstd[:, None]
    y = (xhat * gamma + beta).to(x.dtype)
    y2 = (xhat * gamma2 + beta2).to(x.dtype) if gamma2 is not None else None
    return y, y2, mean, rstd


def layernorm_bwd(x, mean, rstd, gamma, dy, dgamma, dbeta, gamma2=None, dy2=None, dgamma2=None, dbeta2=None, dres=None):
    xhat = (x.float() - mean[:, None]) * rstd[:, None]
    d1 = dy.float()
    dgamma += (d1 * xhat).sum(0)
    dbeta += d1.sum(0)
    dxhat = d1 * gamma
    if gamma2 is not None:
        d2 = dy2.float()
        dgamma2 += (d2 * xhat).sum(0)
        dbeta2 += d2.sum(0)
        dxhat = dxhat + d2 * gamma2
    dx = rstd[:, None] * (dxhat - dxhat.mean(1, keepdim=True) - xhat * (dxhat * xhat).mean(1, keepdim=True))
    if dres is not None:
        dx = dx + dres.float()
    return dx.to(x.dtype)


# ----------------------------------------------------------------------------------------------------- GELU / RoPE
def gelu_fwd(x):
    return _gelu(x.float()).to(x.dtype)


def gelu_bwd(x, dy):
    return (dy.float() * _dgelu(x.float())).to(x.dtype)


def rope_qk_inplace(qkv, cos, sin, B, S, nh, hd, rot, inverse=False):
    x = qkv.view(B, S, nh, 3, hd)
    c = torch.cat([cos[:S], cos[:S]], -1)[None, :, None, :]
    s = torch.cat([sin[:S], sin[:S]], -1)[None, :, None, :]
    if inverse:
        s = -s
    for w in (0, 1):
        t = x[:, :, :, w, :rot].float()
        rh = torch.cat([-t[..., rot // 2:], t[..., : rot // 2]], -1)
        x[:, :, :, w, :rot] = (t * c + rh * s).to(qkv.dtype)
    return qkv


# ----------------------------------------------------------------------------------------------------- Embedding
def embedding_fwd(ids, table):
    assert int(ids.min()) >= 0 and int(ids.max()) < table.shape[0], "embedding: id out of range (the kernel traps)"
    return table[ids.reshape(-1)].contiguous()


def embedding3_fwd(ids0, table0, ids1=None, table1=None, ids2=None, table2=None):
    out = table0[ids0.reshape(-1)].float()
    if table1 is not None:
        out = out + table1[ids1.reshape(-1)].float()
    if table2 is not None:
        out = out + table2[ids2.reshape(-1)].float()
    return out.to(table0.dtype)


def embedding_bwd(ids, dout, dtable, padding_idx=None):
    ids = ids.reshape(-1)
    d = dout.float()
    if padding_idx is not None:
        keep = ids != padding_idx
        ids, d = ids[keep], d[keep]
    dtable.index_add_(0, ids, d)


def roberta_position_ids(ids, pad_id: int):
    mask = (ids != pad_id).to(torch.int64)
    return torch.cumsum(mask, 1) * mask + pad_id


def _keep_mask(seed: int, shape, p: float):
    """STAND-IN for the kernels' counter-based masks: like them a pure function of (seed, element position), so that the host logic
    around them (which seed each site uses, that backward and recomputation reuse the forward's seed, that every step draws new masks)
    behaves the same; NOT the kernels' hash — the masks themselves are checked on the GPU (test_attention_gpu.py, test_roberta_gpu.py)."""
    g = torch.Generator().manual_seed(int(seed) & 0x7FFFFFFFFFFFFFFF)
    return torch.rand(shape, generator=g) >= p


def dropout(x, p: float, seed: int, residual=None, out=None):
    y = x.float()
    if p > 0.0:
        y = y * _keep_mask(seed, x.shape, p) / (1.0 - p)
    if residual is not None:
        y = y + residual.float()
    y = y.to(x.dtype)
    if out is None:
        return y
    out.copy_(y)
    return out


# ----------------------------------------------------------------------------------------------------- Cross entropy
def cross_entropy_(logits, labels, V=None, ignore_index=-100, write_grad=True, grad_scale=None):
    T = logits.shape[0]
    V = logits.shape[1] if V is None else V
    labels = labels.reshape(-1)
    valid = labels != ignore_index
    assert bool(((labels >= 0) & (labels < V))[valid].all()), "cross_entropy: label out of range (the kernel traps)"
    n_valid = int(valid.sum())
    x = logits[:, :V].float()
    lse = torch.logsumexp(x, 1)
    safe = torch.where(valid, labels, torch.zeros_like(labels))
    row_loss = torch.where(valid, lse - x.gather(1, safe[:, None])[:, 0], torch.zeros(T))
    loss = row_loss.sum() / max(n_valid, 1)
    if write_grad:
        s = (float(grad_scale) if grad_scale is not None else 1.0) / max(n_valid, 1)
        g = torch.softmax(x, 1)
        g[torch.arange(T), safe] -= 1.0
        g = g * s * valid[:, None]
        logits.zero_()  # padded columns (>= V) and ignored rows carry a zero gradient
        logits[:, :V] = g.to(logits.dtype)
    return loss.to(F32), torch.tensor([n_valid], dtype=torch.int32)


def colsum_(x, out, out2=None, scale=None):
    inc = x.float().sum(0) * (float(scale) if scale is not None else 1.0)
    out += inc
    if out2 is not None:
        out2 += inc
    return out


# ----------------------------------------------------------------------------------------------------- GEMM
def gemm(A, B, *, a_mn=False, b_mn=False, out=None, out_dtype=None, accumulate=False, bias=None, residual=None, gelu=False,
         alpha=None, aux_out=None, dgelu_in=None, dropout_p: float = 0.0, dropout_seed: int = 0, colsum_out=None):
    assert A.dtype == B.dtype and A.dtype in (torch.bfloat16, torch.float16)
    a = A.float().t() if a_mn else A.float()      # [M, K]
    b = B.float() if b_mn else B.float().t()      # [K, N]
    assert a.shape[1] == b.shape[0], "gemm: reduction dims differ"
    acc = a @ b
    if alpha is not None:
        acc = acc * float(alpha)
    if bias is not None:
        acc = acc + bias
    if gelu:
        if aux_out is not None:
            aux_out.copy_(acc.to(aux_out.dtype))
        acc = _gelu(acc)
    if dgelu_in is not None:
        acc = acc * _dgelu(dgelu_in.float())
        if colsum_out is not None:
            colsum_out += acc.sum(0)
    if dropout_p > 0.0:  # the epilogue's mask is K.dropout's: the same function of (seed, output element)
        assert residual is not None
        acc = acc * _keep_mask(dropout_seed, acc.shape, dropout_p) / (1.0 - dropout_p)
    if residual is not None:
        acc = acc + residual.float()
    if out is None:
        assert not accumulate
        return acc.to(A.dtype if out_dtype is None else out_dtype)
    assert out.shape == acc.shape
    if accumulate:
        out += acc.to(out.dtype)
    else:
        out.copy_(acc.to(out.dtype))
    return out


# ----------------------------------------------------------------------------------------------------- Attention
def _scores(q, k, causal, scale):
    s = torch.einsum("bqhd,bkhd->bhqk", q.float(), k.float()) * scale
    if causal:
        S = q.shape[1]
        s = s.masked_fill(torch.ones(S, S, dtype=torch.bool).triu(1), float("-inf"))
    return s


def attention_fwd(q, k, v, causal, scale=None, dropout_p: float = 0.0, dropout_seed: int = 0):
    D = q.shape[3]
    scale = D ** -0.5 if scale is None else scale
    s = _scores(q, k, causal, scale)
    lse = torch.logsumexp(s, -1)                                  # [B, H, S]
    p = torch.exp(s - lse[..., None])
    if dropout_p > 0.0:  # probabilities dropped after the softmax; lse stays that of the undropped scores
        p = p * _keep_mask(dropout_seed, p.shape, dropout_p) / (1.0 - dropout_p)
    o = torch.einsum("bhqk,bkhd->bqhd", p, v.float()).to(q.dtype).contiguous()
    return o, lse.contiguous()


def attention_bwd(q, k, v, o, lse, d_o, dq, dk, dv, causal, scale=None, dropout_p: float = 0.0, dropout_seed: int = 0):
    D = q.shape[3]
    scale = D ** -0.5 if scale is None else scale
    p = torch.exp(_scores(q, k, causal, scale) - lse[..., None])
    do = d_o.float()
    dp = torch.einsum("bqhd,bkhd->bhqk", do, v.float())
    pd = p
    if dropout_p > 0.0:
        m = _keep_mask(dropout_seed, p.shape, dropout_p) / (1.0 - dropout_p)
        pd, dp = p * m, dp * m
    dv.copy_(torch.einsum("bhqk,bqhd->bkhd", pd, do).to(dv.dtype))
    delta = (do * o.float()).sum(-1).permute(0, 2, 1)             # [B, H, S]; = sum_k p * dp(masked) for the dropped forward too
    ds = p * (dp - delta[..., None]) * scale
    dq.copy_(torch.einsum("bhqk,bkhd->bqhd", ds, k.float()).to(dq.dtype))
    dk.copy_(torch.einsum("bhqk,bqhd->bkhd", ds, q.float()).to(dk.dtype))
    return dq, dk, dv


# ----------------------------------------------------------------------------------------------------- Optimizer
def adam_step(p, g, m, v, p_bf16, state_base, chunk_start, chunk_len, chunk_group, groups, grad_scale=None, zero_grad=False,
              chunk_state=None, skip_flag=None, g_packed=False, p_packed=False):
    """adam.cu: adam_kernel, chunk by chunk."""
    gs = float(grad_scale) if grad_scale is not None else 1.0
    skip = skip_flag is not None and bool(int(skip_flag))
    starts, lens, grps = chunk_start.tolist(), chunk_len.tolist(), chunk_group.tolist()
    soffs = chunk_state.tolist() if chunk_state is not None else [s - state_base for s in starts]
    for start, n, gi, so in zip(starts, lens, grps, soffs):
        gsl = g[so:so + n] if g_packed else g[start:start + n]
        if skip:
            if zero_grad:
                gsl.zero_()
            continue
        h = groups[gi]
        psl = p[so:so + n] if p_packed else p[start:start + n]
        msl, vsl = m[so:so + n], v[so:so + n]
        grad = gsl * gs
        if h["adamw_mode"]:
            psl.mul_(1.0 - h["lr"] * h["weight_decay"])
        else:
            grad = grad + h["weight_decay"] * psl
        msl.mul_(h["beta1"]).add_(grad, alpha=1.0 - h["beta1"])
        vsl.mul_(h["beta2"]).addcmul_(grad, grad, value=1.0 - h["beta2"])
        denom = vsl.sqrt() / math.sqrt(h["bias_corr2"]) + h["eps"]
        psl.addcdiv_(msl, denom, value=-h["lr"] / h["bias_corr1"])
        if zero_grad:
            gsl.zero_()
        if p_bf16 is not None:
            p_bf16[start:start + n] = psl.to(p_bf16.dtype)


def sumsq_(x, out):
    out += (x.double() ** 2).sum().float()
    return out


def sumsq_chunks_(x, chunk_start, chunk_len, out, partials=None):
    tot = torch.zeros((), dtype=torch.float64)
    for s, n in zip(chunk_start.tolist(), chunk_len.tolist()):
        tot += (x[s:s + n].double() ** 2).sum()
    out += tot.float()
    return out


def clip_coef(sumsq, max_norm, loss_scale=None, found_inf=None):
    s = float(loss_scale) if loss_scale is not None else 1.0
    if found_inf is not None:
        found_inf.fill_(0 if bool(torch.isfinite(sumsq)) else 1)
    norm = sumsq.sqrt() / s
    coef = torch.clamp(max_norm / (norm + 1e-6), max=1.0) if max_norm and max_norm > 0 else torch.ones(())
    return norm, (coef / s).to(F32)


def loss_scale_update(scale, growth_tracker, hysteresis_left, found_inf, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000,
                      min_scale=1.0, hysteresis=1):
    """adam.cu: loss_scale_update_kernel."""
    if int(found_inf):
        left = int(hysteresis_left) - 1
        if hysteresis <= 1 or left <= 0:
            scale.fill_(max(float(scale) * backoff_factor, min_scale))
            left = hysteresis if hysteresis <= 1 else 1
        hysteresis_left.fill_(left)
        growth_tracker.zero_()
    else:
        t = int(growth_tracker) + 1
        if t >= growth_interval:
            grown = float(scale) * growth_factor
            if math.isfinite(grown):
                scale.fill_(grown)
            growth_tracker.zero_()
            hysteresis_left.fill_(hysteresis)
        else:
            growth_tracker.fill_(t)


def cast_f32_to_bf16(src, dst):
    dst.copy_(src)


NAMES = ["layernorm_fwd", "layernorm_bwd", "gelu_fwd", "gelu_bwd", "rope_qk_inplace", "embedding_fwd", "embedding3_fwd", "embedding_bwd",
         "roberta_position_ids", "dropout", "cross_entropy_", "colsum_", "gemm", "attention_fwd", "attention_bwd", "adam_step", "sumsq_",
         "sumsq_chunks_", "clip_coef", "loss_scale_update", "cast_f32_to_bf16"]


def install(monkeypatch=None):
    """Replace the C-ABI wrappers of `multimodal_llm_pretraining_b200.kernels` by the statements above (test process only).
    With pytest's `monkeypatch` the originals come back at the end of the test; without it (spawned gloo workers) the process ends."""
    import multimodal_llm_pretraining_b200.kernels as K

    g = globals()
    for n in NAMES:
        assert hasattr(K, n), n  # a renamed wrapper must not be silently skipped
        if monkeypatch is not None:
            monkeypatch.setattr(K, n, g[n])
        else:
            setattr(K, n, g[n])
    return K
