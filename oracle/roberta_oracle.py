"""ORACLE — test infrastructure only (never imported by the product path).

CPU restatement, in plain PyTorch fp32 tensor ops, of the arithmetic the reference's RoBERTa path executes for one
masked-LM pretraining step (the model src/models/roberta.py:15-18 builds: transformers.RobertaForMaskedLM, eager attention).
The arithmetic lives in transformers (pinned 4.47.1, /root/reference/pyproject.toml:16; 5.5.0 in this image, same math).
"HF:" = site-packages/transformers/.  The forward is restated op by op; gradients come from torch.autograd over this
restatement (so they are independent of HF's module code, not of torch's differentiation rules).

Parity pin: tests/golden/roberta_tiny.pt is produced by tests/golden/make_golden_roberta.py from the REAL
transformers.RobertaForMaskedLM (dropout 0) + torch.optim.Adam; tests/test_oracle.py checks this file against it.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file.
"""

from __future__ import annotations

import math

import torch


def position_ids(ids: torch.Tensor, pad: int) -> torch.Tensor:
    """HF:models/roberta/modeling_roberta.py:146-159 (create_position_ids_from_input_ids)."""
    mask = (ids != pad).to(torch.int64)
    return torch.cumsum(mask, dim=1) * mask + pad


def layer_norm(x, w, b, eps):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def gelu(x):
    """HF:activations.py GELUActivation (exact erf)."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def roberta_logits(P: dict, ids: torch.Tensor, cfg: dict) -> torch.Tensor:
    """P: HF-named fp32 parameters. ids int64 [B,S]. Returns logits [B,S,V] (dropout = 0)."""
    h, nh, L, eps = cfg["hidden_size"], cfg["num_attention_heads"], cfg["num_hidden_layers"], cfg["layer_norm_eps"]
    hd = h // nh
    B, S = ids.shape
    e = "roberta.embeddings"
    # HF:modeling_roberta.py:79-144: word + token_type(0) + position, LayerNorm
    # both tables are nn.Embedding(padding_idx=pad): the padding row receives no gradient from the lookup (:61,74-76)
    pad = cfg["pad_token_id"]
    emb = torch.nn.functional.embedding
    x = emb(ids, P[f"{e}.word_embeddings.weight"], padding_idx=pad) + P[f"{e}.token_type_embeddings.weight"][0] + \
        emb(position_ids(ids, pad), P[f"{e}.position_embeddings.weight"], padding_idx=pad)
    x = layer_norm(x, P[f"{e}.LayerNorm.weight"], P[f"{e}.LayerNorm.bias"], eps)
    for i in range(L):
        p = f"roberta.encoder.layer.{i}"
        # HF:modeling_roberta.py:190-254 + eager_attention_forward :162-187
        def lin(t, n):
            return t @ P[f"{p}.{n}.weight"].t() + P[f"{p}.{n}.bias"]
        q = lin(x, "attention.self.query").view(B, S, nh, hd).transpose(1, 2)
        k = lin(x, "attention.self.key").view(B, S, nh, hd).transpose(1, 2)
        v = lin(x, "attention.self.value").view(B, S, nh, hd).transpose(1, 2)
        att = torch.softmax((q @ k.transpose(-1, -2)) * hd ** -0.5, dim=-1)
        ctx = (att @ v).transpose(1, 2).reshape(B, S, h)
        # HF:modeling_roberta.py:334-347 (RobertaSelfOutput): LN(dense(ctx) + x)
        x = layer_norm(lin(ctx, "attention.output.dense") + x, P[f"{p}.attention.output.LayerNorm.weight"], P[f"{p}.attention.output.LayerNorm.bias"], eps)
        # HF:modeling_roberta.py:377-404 (Intermediate + Output)
        inter = gelu(lin(x, "intermediate.dense"))
        x = layer_norm(lin(inter, "output.dense") + x, P[f"{p}.output.LayerNorm.weight"], P[f"{p}.output.LayerNorm.bias"], eps)
    # HF:modeling_roberta.py:882-901 (RobertaLMHead), decoder tied to the word embeddings (:798-801)
    d = gelu(x @ P["lm_head.dense.weight"].t() + P["lm_head.dense.bias"])
    d = layer_norm(d, P["lm_head.layer_norm.weight"], P["lm_head.layer_norm.bias"], eps)
    return d @ P[f"{e}.word_embeddings.weight"].t() + P["lm_head.bias"]


def roberta_loss_and_grads(P: dict, ids: torch.Tensor, labels: torch.Tensor, cfg: dict):
    """Mean CE over all positions with label != -100 (HF:modeling_roberta.py:867-871, CrossEntropyLoss)."""
    Pg = {k: v.detach().clone().float().requires_grad_(True) for k, v in P.items() if not k.startswith("lm_head.decoder.")}
    logits = roberta_logits(Pg, ids, cfg)
    lp = torch.log_softmax(logits.reshape(-1, logits.shape[-1]), dim=-1)
    lab = labels.reshape(-1)
    valid = lab != -100
    loss = -(lp[torch.arange(lab.numel()), lab.clamp_min(0)] * valid).sum() / valid.sum().clamp_min(1)
    loss.backward()
    return loss.detach(), {k: v.grad.detach() for k, v in Pg.items()}, logits.detach()
