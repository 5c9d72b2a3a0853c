"""ORACLE — test infrastructure only (never imported by the product path).

CPU restatement, in plain PyTorch fp32/fp64 tensor ops, of the arithmetic the reference's hot path executes for a
Pythia / GPT-NeoX pretraining step. The reference tree contains none of this arithmetic itself: it lives
in the third-party dependency transformers (pinned 4.47.1 in /root/reference/pyproject.toml:16; 5.5.0 is what this image
has, the math of these classes is unchanged) and torch.optim.Adam.  Each function cites the file:line it follows
("HF:" = site-packages/transformers/).

Parity pin: tests/golden/neox_tiny_*.pt are produced by tests/golden/make_golden.py, which runs the REAL
transformers.GPTNeoXForCausalLM + torch.optim.Adam (the reference's own call path: src/models/pythia.py:15-22,43-67;
src/benchmarking/utils.py:61-80) on seeded inputs; tests/test_oracle.py checks this restatement against those vectors
(and, where transformers is importable, against the live HF module).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this file.
"""

from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def rotary_tables(S: int, rot: int, base: float = 10000.0, dtype=torch.float32):
    """HF:models/gpt_neox/modeling_gpt_neox.py:52-116 (GPTNeoXRotaryEmbedding): inv_freq = base^(-2i/rot), fp32 angles,
    emb = cat(freqs, freqs)."""
    inv_freq = 1.0 / (base ** (torch.arange(0, rot, 2, dtype=torch.int64).float() / rot))
    freqs = torch.arange(S, dtype=torch.float32)[:, None] * inv_freq[None, :]
    emb = torch.cat([freqs, freqs], dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def rotate_half(x):
    """HF:modeling_gpt_neox.py:119-123."""
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def apply_rotary(q, k, cos, sin):
    """HF:modeling_gpt_neox.py:126-159: only the first `rot` dims rotate; q,k are [B, nh, S, hd]; cos/sin [S, rot]."""
    rot = cos.shape[-1]
    cos, sin = cos[None, None], sin[None, None]
    q_rot, q_pass = q[..., :rot], q[..., rot:]
    k_rot, k_pass = k[..., :rot], k[..., rot:]
    q_emb = torch.cat([q_rot * cos + rotate_half(q_rot) * sin, q_pass], dim=-1)
    k_emb = torch.cat([k_rot * cos + rotate_half(k_rot) * sin, k_pass], dim=-1)
    return q_emb, k_emb


def gelu_erf(x):
    """HF:activations.py GELUActivation (exact erf)."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def neox_layer(x, P: dict, prefix: str, nh: int, rot: int, eps: float, base: float):
    """HF:modeling_gpt_neox.py:258-289 (GPTNeoXLayer, use_parallel_residual=True) + :203-244 (attention) + :38-49 (MLP)."""
    B, S, h = x.shape
    hd = h // nh
    a = F.layer_norm(x, (h,), P[f"{prefix}.input_layernorm.weight"], P[f"{prefix}.input_layernorm.bias"], eps)
    qkv = a @ P[f"{prefix}.attention.query_key_value.weight"].T + P[f"{prefix}.attention.query_key_value.bias"]
    qkv = qkv.view(B, S, nh, 3 * hd).transpose(1, 2)  # per-head interleaved [q|k|v] (:211-215)
    q, k, v = qkv.chunk(3, dim=-1)
    cos, sin = rotary_tables(S, rot, base, x.dtype)
    q, k = apply_rotary(q, k, cos, sin)
    s = (q @ k.transpose(-1, -2)) * hd ** -0.5
    mask = torch.ones(S, S, dtype=torch.bool).tril()
    s = s.masked_fill(~mask, float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, S, h)
    att = o @ P[f"{prefix}.attention.dense.weight"].T + P[f"{prefix}.attention.dense.bias"]
    m_in = F.layer_norm(x, (h,), P[f"{prefix}.post_attention_layernorm.weight"], P[f"{prefix}.post_attention_layernorm.bias"], eps)
    m = gelu_erf(m_in @ P[f"{prefix}.mlp.dense_h_to_4h.weight"].T + P[f"{prefix}.mlp.dense_h_to_4h.bias"])
    m = m @ P[f"{prefix}.mlp.dense_4h_to_h.weight"].T + P[f"{prefix}.mlp.dense_4h_to_h.bias"]
    return m + att + x  # :279-282


def neox_logits(P: dict, input_ids, cfg: dict):
    """HF:modeling_gpt_neox.py:331-381 (GPTNeoXModel.forward) + :464 (embed_out)."""
    x = P["gpt_neox.embed_in.weight"][input_ids]
    hd = cfg["hidden_size"] // cfg["num_attention_heads"]
    rot = int(hd * cfg.get("rotary_pct", 0.25))
    for i in range(cfg["num_hidden_layers"]):
        x = neox_layer(x, P, f"gpt_neox.layers.{i}", cfg["num_attention_heads"], rot, cfg.get("layer_norm_eps", 1e-5),
                       cfg.get("rotary_emb_base", 10000.0))
    h = cfg["hidden_size"]
    x = F.layer_norm(x, (h,), P["gpt_neox.final_layer_norm.weight"], P["gpt_neox.final_layer_norm.bias"], cfg.get("layer_norm_eps", 1e-5))
    return x @ P["embed_out.weight"].T


def causal_lm_loss(logits, labels, ignore_index: int = -100):
    """HF:loss/loss_utils.py:45-67 (ForCausalLMLoss): upcast to fp32, pad labels with -100 and shift, mean CE."""
    logits = logits.float()
    labels = F.pad(labels, (0, 1), value=ignore_index)
    shift = labels[..., 1:].contiguous()
    return F.cross_entropy(logits.view(-1, logits.shape[-1]), shift.view(-1), ignore_index=ignore_index, reduction="mean")


def neox_loss(P: dict, input_ids, labels, cfg: dict):
    return causal_lm_loss(neox_logits(P, input_ids, cfg), labels)


def neox_loss_and_grads(P: dict, input_ids, labels, cfg: dict):
    """Forward + backward (torch autograd over the restated ops). Returns (loss, {name: grad})."""
    Q = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
    loss = neox_loss(Q, input_ids, labels, cfg)
    loss.backward()
    return loss.detach(), {k: v.grad for k, v in Q.items()}


def clip_coef(grads: dict, max_norm: float):
    """torch.nn.utils.clip_grad_norm_ as called by src/benchmarking/utils.py:66-70."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0) if max_norm > 0 else torch.tensor(1.0)
    return total, coef


def adam_step(P: dict, grads: dict, state: dict, *, lr: float, betas=(0.9, 0.95), eps: float = 1e-8, weight_decay: float = 0.0,
              adamw: bool = False):
    """torch.optim.Adam (L2-coupled; what src/models/pythia.py:43-45 selects) / AdamW single-tensor update, in place."""
    state["step"] = state.get("step", 0) + 1
    t = state["step"]
    b1, b2 = betas
    for k, p in P.items():
        g = grads[k]
        if adamw:
            p.mul_(1 - lr * weight_decay)
        elif weight_decay != 0:
            g = g + weight_decay * p
        m = state.setdefault(("m", k), torch.zeros_like(p))
        v = state.setdefault(("v", k), torch.zeros_like(p))
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = v.sqrt() / math.sqrt(1 - b2 ** t) + eps
        p.addcdiv_(m, denom, value=-lr / (1 - b1 ** t))


def cosine_with_min_lr(step: int, warmup: int, total: int, min_lr_rate: float, num_cycles: float = 0.5) -> float:
    """HF:optimization.py:324-334."""
    if step < warmup:
        return step / max(1, warmup)
    progress = (step - warmup) / max(1, total - warmup)
    f = 0.5 * (1.0 + math.cos(math.pi * num_cycles * 2.0 * progress))
    return max(0.0, f * (1 - min_lr_rate) + min_lr_rate)


def train_steps(P: dict, batches, cfg: dict, *, lr: float, betas=(0.9, 0.95), eps=1e-8, max_grad_norm: float = 1.0,
                warmup: int = 0, total_steps: int = 1, min_lr_rate: float = 0.1):
    """The reference step: fwd+bwd (manual_training_step) then clip -> Adam -> scheduler -> zero_grad
    (manual_optimization_step, src/benchmarking/utils.py:61-80). LR for step t uses the scheduler value at t (HF calls
    scheduler.step() after optimizer.step(), initial lr = base * lambda(0))."""
    state: dict = {}
    losses = []
    for t, ids in enumerate(batches):
        loss, grads = neox_loss_and_grads(P, ids, ids, cfg)
        _, coef = clip_coef(grads, max_grad_norm)
        grads = {k: g * coef for k, g in grads.items()}
        cur_lr = lr * (cosine_with_min_lr(t, warmup, total_steps, min_lr_rate) if warmup or total_steps > 1 else 1.0)
        adam_step(P, grads, state, lr=cur_lr, betas=betas, eps=eps)
        losses.append(loss.item())
    return losses
