"""B200-native drop-in for transformers.RobertaForMaskedLM (the module `src/models/roberta.py:15-18` builds).

Same parameter names / shapes / state_dict keys as HF (HF:models/roberta/modeling_roberta.py:56-144,190-254,334-468,797-901,
including the tied `lm_head.decoder.weight` = word embeddings and `lm_head.decoder.bias` = `lm_head.bias`), same call
convention `model(input_ids=..., labels=...)["loss"]`. The post-LN block reuses the Pythia path's kernels with
bidirectional attention: one fused QKV GEMM (query/key/value weights are contiguous in the flat store, so the three HF
matrices ARE one [3h, h] operand), tcgen05 flash attention reading q|k|v in place, GEMM epilogues for bias / bias+GELU /
bias+residual, LayerNorm, in-place cross entropy over all positions.

Layout notes
  * V = 50265 is odd: the word-embedding / decoder matrix and the decoder bias are allocated with zero padding to 50304 rows
    (FlatParams alloc_shape) so that logits rows are 16-byte aligned; the padding is never updated and never scored.
  * Dropout (hidden 0.1, attention-probability 0.1 in roberta-large): both are counter-based masks recomputed in backward
    (b200_dropout for hidden states; inside the flash-attention kernels for the softmax probabilities). p = 0 is the
    bit-comparable parity configuration; with p > 0 parity is checked against a reference that applies the same masks.
"""

from __future__ import annotations

from types import SimpleNamespace

import torch
from torch import nn

from . import kernels as K
from .flat import FlatParams
from .modeling_gpt_neox import BF16, ModelOutput, _FlatModule, _Params

V_ALIGN = 64


def _pad_vocab(V: int) -> int:
    return (V + V_ALIGN - 1) // V_ALIGN * V_ALIGN


def roberta_param_shapes(cfg) -> list[tuple]:
    h, V, L, I = cfg.hidden_size, cfg.vocab_size, cfg.num_hidden_layers, cfg.intermediate_size
    Vp = _pad_vocab(V)
    e = "roberta.embeddings"
    shapes: list[tuple] = [
        (f"{e}.word_embeddings.weight", (V, h), (Vp, h)),
        (f"{e}.position_embeddings.weight", (cfg.max_position_embeddings, h)),
        (f"{e}.token_type_embeddings.weight", (cfg.type_vocab_size, h)),
        (f"{e}.LayerNorm.weight", (h,)), (f"{e}.LayerNorm.bias", (h,)),
    ]
    for i in range(L):
        p = f"roberta.encoder.layer.{i}"
        shapes += [
            # query | key | value back to back: one [3h, h] GEMM operand, one [3h] bias
            (f"{p}.attention.self.query.weight", (h, h)), (f"{p}.attention.self.key.weight", (h, h)),
            (f"{p}.attention.self.value.weight", (h, h)),
            (f"{p}.attention.self.query.bias", (h,)), (f"{p}.attention.self.key.bias", (h,)), (f"{p}.attention.self.value.bias", (h,)),
            (f"{p}.attention.output.dense.weight", (h, h)), (f"{p}.attention.output.dense.bias", (h,)),
            (f"{p}.attention.output.LayerNorm.weight", (h,)), (f"{p}.attention.output.LayerNorm.bias", (h,)),
            (f"{p}.intermediate.dense.weight", (I, h)), (f"{p}.intermediate.dense.bias", (I,)),
            (f"{p}.output.dense.weight", (h, I)), (f"{p}.output.dense.bias", (h,)),
            (f"{p}.output.LayerNorm.weight", (h,)), (f"{p}.output.LayerNorm.bias", (h,)),
        ]
    shapes += [
        ("lm_head.bias", (V,), (Vp,)),
        ("lm_head.dense.weight", (h, h)), ("lm_head.dense.bias", (h,)),
        ("lm_head.layer_norm.weight", (h,)), ("lm_head.layer_norm.bias", (h,)),
    ]
    return shapes


class _TiedDecoder(nn.Module):
    """`lm_head.decoder`: its weight / bias are the word-embedding matrix and `lm_head.bias` (HF _tied_weights_keys,
    modeling_roberta.py:798-801). Registered as plain attributes so named_parameters() lists each tensor once while
    state_dict() still carries HF's four keys (see B200RobertaForMaskedLM.state_dict)."""

    def __init__(self, weight: nn.Parameter, bias: nn.Parameter):
        super().__init__()
        object.__setattr__(self, "weight", weight)
        object.__setattr__(self, "bias", bias)


class B200RobertaForMaskedLM(_FlatModule):
    supports_gradient_checkpointing = True
    main_input_name = "input_ids"

    def __init__(self, config, allow_missing_attention_dropout: bool = False):
        super().__init__()
        self.config = cfg = config
        assert getattr(cfg, "hidden_act", "gelu") == "gelu", "only exact-erf GELU is built"
        assert getattr(cfg, "position_embedding_type", "absolute") == "absolute"
        self.h, self.nh = cfg.hidden_size, cfg.num_attention_heads
        self.hd = self.h // self.nh
        self.L, self.V, self.Vp = cfg.num_hidden_layers, cfg.vocab_size, _pad_vocab(cfg.vocab_size)
        self.inter = cfg.intermediate_size
        self.eps = cfg.layer_norm_eps
        self.pad_id = cfg.pad_token_id
        self.p_hidden = float(getattr(cfg, "hidden_dropout_prob", 0.0))
        self.p_attn = float(getattr(cfg, "attention_probs_dropout_prob", 0.0))
        self.allow_missing_attention_dropout = allow_missing_attention_dropout
        if self.hd not in (64, 128, 256):
            raise NotImplementedError(f"head_dim {self.hd} has no tcgen05 attention kernel (built: 64, 128, 256)")

        self.flat = f = FlatParams(roberta_param_shapes(cfg))
        body = nn.Module()
        emb = nn.Module()
        e = "roberta.embeddings"
        emb.word_embeddings = _Params(f, f"{e}.word_embeddings", ("weight",))
        emb.position_embeddings = _Params(f, f"{e}.position_embeddings", ("weight",))
        emb.token_type_embeddings = _Params(f, f"{e}.token_type_embeddings", ("weight",))
        emb.LayerNorm = _Params(f, f"{e}.LayerNorm", ("weight", "bias"))
        body.embeddings = emb
        enc = nn.Module()
        layers = []
        for i in range(self.L):
            p = f"roberta.encoder.layer.{i}"
            lyr = nn.Module()
            att = nn.Module()
            slf = nn.Module()
            slf.query = _Params(f, f"{p}.attention.self.query", ("weight",))
            slf.key = _Params(f, f"{p}.attention.self.key", ("weight",))
            slf.value = _Params(f, f"{p}.attention.self.value", ("weight",))
            for nm in ("query", "key", "value"):
                getattr(slf, nm).register_parameter("bias", f.make_parameter(f"{p}.attention.self.{nm}.bias"))
            att.self = slf
            out = nn.Module()
            out.dense = _Params(f, f"{p}.attention.output.dense", ("weight", "bias"))
            out.LayerNorm = _Params(f, f"{p}.attention.output.LayerNorm", ("weight", "bias"))
            att.output = out
            lyr.attention = att
            inter = nn.Module()
            inter.dense = _Params(f, f"{p}.intermediate.dense", ("weight", "bias"))
            lyr.intermediate = inter
            o2 = nn.Module()
            o2.dense = _Params(f, f"{p}.output.dense", ("weight", "bias"))
            o2.LayerNorm = _Params(f, f"{p}.output.LayerNorm", ("weight", "bias"))
            lyr.output = o2
            layers.append(lyr)
        enc.layer = nn.ModuleList(layers)
        body.encoder = enc
        self.roberta = body
        head = nn.Module()
        head.register_parameter("bias", f.make_parameter("lm_head.bias"))
        head.dense = _Params(f, "lm_head.dense", ("weight", "bias"))
        head.layer_norm = _Params(f, "lm_head.layer_norm", ("weight", "bias"))
        head.decoder = _TiedDecoder(emb.word_embeddings.weight, head.bias)
        self.lm_head = head
        self.grad_ready_hook = None
        self.param_wait_hook = None
        self._fwd_buckets = list(reversed(self.comm_buckets()))  # embeddings, layers 0..L-1, head
        self.gradient_checkpointing = False
        self._step_seed = 0  # bumped once per training forward: every micro-batch draws fresh masks
        self._cur_seed = 0   # the seed base of the forward/backward currently running
        self.reset_parameters()

    # ------------------------------------------------------------------ init / HF surface
    @torch.no_grad()
    def reset_parameters(self, generator: torch.Generator | None = None) -> None:
        """HF _init_weights: Linear / Embedding N(0, initializer_range) (padding_idx rows zeroed), biases 0, LayerNorm (1, 0)."""
        std = getattr(self.config, "initializer_range", 0.02)
        for name, p in self.named_parameters():
            if name.endswith("LayerNorm.weight") or name.endswith("layer_norm.weight"):
                p.fill_(1.0)
            elif name.endswith(".bias") or name == "lm_head.bias":
                p.zero_()
            else:
                p.normal_(0.0, std, generator=generator)
        if self.pad_id is not None:
            self.roberta.embeddings.word_embeddings.weight[self.pad_id].zero_()
            self.roberta.embeddings.position_embeddings.weight[self.pad_id].zero_()

    def state_dict(self, *args, **kwargs):
        sd = super().state_dict(*args, **kwargs)
        prefix = kwargs.get("prefix", args[1] if len(args) > 1 else "")
        sd[f"{prefix}lm_head.decoder.weight"] = sd[f"{prefix}roberta.embeddings.word_embeddings.weight"]
        sd[f"{prefix}lm_head.decoder.bias"] = sd[f"{prefix}lm_head.bias"]
        return sd

    def load_hf_state_dict(self, sd: dict[str, torch.Tensor]) -> None:
        own = nn.Module.state_dict(self)
        missing = [k for k in own if k not in sd]
        if missing:
            raise KeyError(f"missing keys: {missing[:4]}…")
        with torch.no_grad():
            for k, v in own.items():
                v.copy_(sd[k].to(v.dtype))

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        sd = {k: v for k, v in state_dict.items() if not k.startswith("lm_head.decoder.")}
        return super().load_state_dict(sd, strict=strict, assign=assign)

    def gradient_checkpointing_enable(self, gradient_checkpointing_kwargs=None) -> None:
        """Per-layer activation checkpointing (src/train.py:112): keep only each layer's input and recompute the layer in
        backward. The dropout masks are pure functions of (step seed, site, element), so the recomputation reproduces them."""
        self.gradient_checkpointing = True

    def gradient_checkpointing_disable(self) -> None:
        self.gradient_checkpointing = False

    @property
    def is_gradient_checkpointing(self) -> bool:
        return self.gradient_checkpointing

    @property
    def device(self) -> torch.device:
        return self.flat.device

    @property
    def dtype(self) -> torch.dtype:
        return torch.float32

    def get_input_embeddings(self):
        return self.roberta.embeddings.word_embeddings

    def get_output_embeddings(self):
        return self.lm_head.decoder

    def num_parameters(self, only_trainable: bool = False) -> int:
        return sum(p._b200_flat[2] for p in self.parameters())

    # ------------------------------------------------------------------ helpers
    def _w(self, name):
        return self.flat.wview(name)

    def _p(self, name):
        return self.flat.pview(name)

    def _g(self, name):
        return self.flat.gview(name)

    def _qkv_w(self, p):  # [3h, h] bf16: query | key | value rows
        return self.flat.wview_span(f"{p}.attention.self.query.weight", (3 * self.h, self.h))

    def _qkv_b(self, p, buf):
        return self.flat.view_span(buf, f"{p}.attention.self.query.bias", (3 * self.h,))

    def _qkv_gw(self, p):
        return self.flat.gview_span(f"{p}.attention.self.query.weight", (3 * self.h, self.h))

    def _qkv_gb(self, p):
        return self.flat.gview_span(f"{p}.attention.self.query.bias", (3 * self.h,))

    def _drop(self, x, residual, site: int):
        """dropout(x) + residual with the mask of (step seed, site); identity add when p == 0 is done by the GEMM epilogue."""
        return K.dropout(x, self.p_hidden, self._seed(site), residual=residual)

    def _seed(self, site: int) -> int:
        return (self._cur_seed * 1_000_003 + site) & 0xFFFFFFFFFFFF

    supports_weight_sharding = True

    def weight_bucket_deps(self) -> dict[tuple[int, int], list[tuple[int, int]]]:
        """ZeRO-3: buckets that must be resident TOGETHER with a bucket — the head reads the tied decoder matrix, which lives in the
        embeddings bucket."""
        b = self.comm_buckets()
        return {tuple(b[0]): [tuple(b[-1])]}

    def comm_buckets(self) -> list[tuple[int, int]]:
        """Flat-grad ranges in backward completion order: head, layers L-1..0, embeddings (the tied decoder gradient is
        only final after the embedding backward, so the embeddings bucket comes last and includes it)."""
        f = self.flat
        head = f.range_of(["lm_head.bias", "lm_head.dense.weight", "lm_head.dense.bias", "lm_head.layer_norm.weight", "lm_head.layer_norm.bias"])
        layers = [f.range_of([n for n in f.names if n.startswith(f"roberta.encoder.layer.{i}.")]) for i in reversed(range(self.L))]
        emb = f.range_of([n for n in f.names if n.startswith("roberta.embeddings.")])
        return [head] + layers + [emb]

    # ------------------------------------------------------------------ forward / backward of one layer
    def _layer_fwd(self, i: int, x: torch.Tensor, B: int, S: int, train: bool):
        p = f"roberta.encoder.layer.{i}"
        h, nh, hd = self.h, self.nh, self.hd
        drop = train and self.p_hidden > 0.0
        qkv = K.gemm(x, self._qkv_w(p), bias=self.flat.pview_span(f"{p}.attention.self.query.bias", (3 * self.h,)))  # [T, 3h] = q | k | v
        q4 = qkv.view(B, S, 3, nh, hd)
        pa = self.p_attn if train else 0.0
        o, lse = K.attention_fwd(q4[:, :, 0], q4[:, :, 1], q4[:, :, 2], causal=False, scale=hd ** -0.5, dropout_p=pa, dropout_seed=self._seed(4 * i + 3))
        o2 = o.view(B * S, h)
        wo, bo = self._w(f"{p}.attention.output.dense.weight"), self._p(f"{p}.attention.output.dense.bias")
        # hidden dropout fused into the GEMM epilogue: dropout(W o + b) + x in one pass (same mask as K.dropout(seed), which backward
        # applies to the gradient)
        s1 = K.gemm(o2, wo, bias=bo, residual=x, dropout_p=self.p_hidden if drop else 0.0, dropout_seed=self._seed(4 * i + 1))
        x1, _, mean1, rstd1 = K.layernorm_fwd(s1, self._p(f"{p}.attention.output.LayerNorm.weight"), self._p(f"{p}.attention.output.LayerNorm.bias"), self.eps)
        h1 = torch.empty(B * S, self.inter, dtype=x.dtype, device=x.device)
        g = K.gemm(x1, self._w(f"{p}.intermediate.dense.weight"), bias=self._p(f"{p}.intermediate.dense.bias"), gelu=True, aux_out=h1)
        w2, b2 = self._w(f"{p}.output.dense.weight"), self._p(f"{p}.output.dense.bias")
        s2 = K.gemm(g, w2, bias=b2, residual=x1, dropout_p=self.p_hidden if drop else 0.0, dropout_seed=self._seed(4 * i + 2))
        x2, _, mean2, rstd2 = K.layernorm_fwd(s2, self._p(f"{p}.output.LayerNorm.weight"), self._p(f"{p}.output.LayerNorm.bias"), self.eps)
        return x2, (x, qkv, o, lse, s1, mean1, rstd1, x1, h1, g, s2, mean2, rstd2)

    def _layer_bwd(self, i: int, saved, dx2: torch.Tensor, B: int, S: int, train: bool) -> torch.Tensor:
        p = f"roberta.encoder.layer.{i}"
        x, qkv, o, lse, s1, mean1, rstd1, x1, h1, g, s2, mean2, rstd2 = saved
        h, nh, hd = self.h, self.nh, self.hd
        drop = train and self.p_hidden > 0.0
        # x2 = LN2(s2), s2 = drop(z) + x1, z = W2 g + b2
        ds2 = K.layernorm_bwd(s2, mean2, rstd2, self._p(f"{p}.output.LayerNorm.weight"), dx2,
                              self._g(f"{p}.output.LayerNorm.weight"), self._g(f"{p}.output.LayerNorm.bias"))
        dz = K.dropout(ds2, self.p_hidden, self._seed(4 * i + 2)) if drop else ds2
        K.gemm(dz, g, a_mn=True, b_mn=True, out=self._g(f"{p}.output.dense.weight"), accumulate=True)
        K.colsum_(dz, self._g(f"{p}.output.dense.bias"))
        from .modeling_gpt_neox import FUSED_BIAS_GRAD  # opt-in, see there

        dh1 = K.gemm(dz, self._w(f"{p}.output.dense.weight"), b_mn=True, dgelu_in=h1,
                     colsum_out=self._g(f"{p}.intermediate.dense.bias") if FUSED_BIAS_GRAD else None)
        K.gemm(dh1, x1, a_mn=True, b_mn=True, out=self._g(f"{p}.intermediate.dense.weight"), accumulate=True)
        if not FUSED_BIAS_GRAD:
            K.colsum_(dh1, self._g(f"{p}.intermediate.dense.bias"))
        dx1 = K.gemm(dh1, self._w(f"{p}.intermediate.dense.weight"), b_mn=True, residual=ds2)  # + residual branch of s2
        # x1 = LN1(s1), s1 = drop(y) + x, y = Wo ctx + bo
        ds1 = K.layernorm_bwd(s1, mean1, rstd1, self._p(f"{p}.attention.output.LayerNorm.weight"), dx1,
                              self._g(f"{p}.attention.output.LayerNorm.weight"), self._g(f"{p}.attention.output.LayerNorm.bias"))
        dy = K.dropout(ds1, self.p_hidden, self._seed(4 * i + 1)) if drop else ds1
        o2 = o.view(B * S, h)
        K.gemm(dy, o2, a_mn=True, b_mn=True, out=self._g(f"{p}.attention.output.dense.weight"), accumulate=True)
        K.colsum_(dy, self._g(f"{p}.attention.output.dense.bias"))
        d_o = K.gemm(dy, self._w(f"{p}.attention.output.dense.weight"), b_mn=True)
        dqkv = torch.empty_like(qkv)
        q4, d4 = qkv.view(B, S, 3, nh, hd), dqkv.view(B, S, 3, nh, hd)
        K.attention_bwd(q4[:, :, 0], q4[:, :, 1], q4[:, :, 2], o, lse, d_o.view(B, S, nh, hd),
                        d4[:, :, 0], d4[:, :, 1], d4[:, :, 2], causal=False, scale=hd ** -0.5,
                        dropout_p=self.p_attn if train else 0.0, dropout_seed=self._seed(4 * i + 3))
        K.gemm(dqkv, x, a_mn=True, b_mn=True, out=self._qkv_gw(p), accumulate=True)
        K.colsum_(dqkv, self._qkv_gb(p))
        return K.gemm(dqkv, self._qkv_w(p), b_mn=True, residual=ds1)  # + residual branch of s1

    # ------------------------------------------------------------------ whole model
    def _embed(self, ids: torch.Tensor, train: bool):
        e = "roberta.embeddings"
        B, S = ids.shape
        self._wait_bucket(self._fwd_buckets[0])  # embeddings (+ the tied decoder matrix the head reads)
        pos = K.roberta_position_ids(ids, self.pad_id)
        tok = torch.zeros_like(ids)  # token_type_ids default to 0 (HF:modeling_roberta.py:106-116)
        emb = K.embedding3_fwd(ids.reshape(-1), self._w(f"{e}.word_embeddings.weight"), pos.reshape(-1), self._w(f"{e}.position_embeddings.weight"),
                               tok.reshape(-1), self._w(f"{e}.token_type_embeddings.weight"))
        x0, _, mean, rstd = K.layernorm_fwd(emb, self._p(f"{e}.LayerNorm.weight"), self._p(f"{e}.LayerNorm.bias"), self.eps)
        if train and self.p_hidden > 0.0:
            x0 = K.dropout(x0, self.p_hidden, self._seed(0))
        return x0, (pos, emb, mean, rstd)

    def _head_logits(self, x: torch.Tensor, keep: bool):
        self._wait_bucket(self._fwd_buckets[-1])
        d_pre = torch.empty_like(x) if keep else None
        d = K.gemm(x, self._w("lm_head.dense.weight"), bias=self._p("lm_head.dense.bias"), gelu=True, aux_out=d_pre)
        n, _, mean, rstd = K.layernorm_fwd(d, self._p("lm_head.layer_norm.weight"), self._p("lm_head.layer_norm.bias"), self.eps)
        w_dec = self.flat.wview_alloc("roberta.embeddings.word_embeddings.weight")  # [Vp, h], zero padding
        b_dec = self.flat.pview_alloc("lm_head.bias")
        logits = K.gemm(n, w_dec, bias=b_dec)  # [T, Vp] bf16
        return logits, (x, d_pre, d, n, mean, rstd)

    def _train_forward(self, ids: torch.Tensor, labels: torch.Tensor):
        B, S = ids.shape
        x, emb_saved = self._embed(ids, True)
        saved_layers = []
        for i in range(self.L):
            self._wait_bucket(self._fwd_buckets[1 + i])
            if self.gradient_checkpointing:
                saved_layers.append(x)  # recompute the layer in backward
                x, _ = self._layer_fwd(i, x, B, S, True)
            else:
                x, sv = self._layer_fwd(i, x, B, S, True)
                saved_layers.append(sv)
        logits, head_saved = self._head_logits(x, keep=True)
        loss, _ = K.cross_entropy_(logits, labels.reshape(-1), V=self.V, write_grad=True, grad_scale=self.loss_scale)
        ctx = SimpleNamespace(ids=ids, B=B, S=S, emb=emb_saved, layers=saved_layers, head=head_saved, dlogits=logits, seed=self._cur_seed)
        return loss, ctx

    def _train_backward(self, ctx, grad_out: torch.Tensor) -> None:
        B, S = ctx.B, ctx.S
        f = self.flat
        self._cur_seed = ctx.seed  # backward recomputes the dropout masks of ITS forward
        alpha = grad_out.reshape(1).to(torch.float32).contiguous()
        hook = self.grad_ready_hook
        buckets = self.comm_buckets()
        self._need_bucket_bwd(buckets[0])  # ZeRO-3: head (+ the tied decoder matrix, see weight_bucket_deps)
        x, d_pre, d, n, mean, rstd = ctx.head
        dl = ctx.dlogits
        g_dec = f.gview_alloc("roberta.embeddings.word_embeddings.weight")
        K.gemm(dl, n, a_mn=True, b_mn=True, out=g_dec, accumulate=True, alpha=alpha)
        # decoder bias grad = alpha * colsum(dlogits)
        K.colsum_(dl, f.gview_alloc("lm_head.bias"), scale=alpha)
        dn = K.gemm(dl, f.wview_alloc("roberta.embeddings.word_embeddings.weight"), b_mn=True, alpha=alpha)
        ctx.dlogits = None
        dd = K.layernorm_bwd(d, mean, rstd, self._p("lm_head.layer_norm.weight"), dn,
                             self._g("lm_head.layer_norm.weight"), self._g("lm_head.layer_norm.bias"))
        dd_pre = K.gelu_bwd(d_pre, dd)
        K.gemm(dd_pre, x, a_mn=True, b_mn=True, out=self._g("lm_head.dense.weight"), accumulate=True)
        K.colsum_(dd_pre, self._g("lm_head.dense.bias"))
        dx = K.gemm(dd_pre, self._w("lm_head.dense.weight"), b_mn=True)
        if hook:
            hook(*buckets[0])
        for i in reversed(range(self.L)):
            sv = ctx.layers[i]
            self._need_bucket_bwd(buckets[1 + (self.L - 1 - i)])
            if isinstance(sv, torch.Tensor):  # checkpointed: recompute this layer's activations (same masks: seed restored above)
                _, sv = self._layer_fwd(i, sv, B, S, True)
            dx = self._layer_bwd(i, sv, dx, B, S, True)
            ctx.layers[i] = None
            if hook:
                hook(*buckets[1 + (self.L - 1 - i)])
        e = "roberta.embeddings"
        pos, emb, mean, rstd = ctx.emb
        if self.p_hidden > 0.0:
            dx = K.dropout(dx, self.p_hidden, self._seed(0))
        demb = K.layernorm_bwd(emb, mean, rstd, self._p(f"{e}.LayerNorm.weight"), dx,
                               self._g(f"{e}.LayerNorm.weight"), self._g(f"{e}.LayerNorm.bias"))
        # both tables are nn.Embedding(padding_idx=pad) in HF: the padding row gets no gradient from the lookup (the tied
        # decoder still contributes to the word-embedding row through the logits GEMM above)
        K.embedding_bwd(ctx.ids.reshape(-1), demb, self._g(f"{e}.word_embeddings.weight"), padding_idx=self.pad_id)
        K.embedding_bwd(pos.reshape(-1), demb, self._g(f"{e}.position_embeddings.weight"), padding_idx=self.pad_id)
        K.colsum_(demb, self._g(f"{e}.token_type_embeddings.weight")[0])  # every token has type 0
        if hook:
            hook(*buckets[-1])

    def forward(self, input_ids: torch.Tensor, labels: torch.Tensor | None = None, attention_mask: torch.Tensor | None = None,
                token_type_ids: torch.Tensor | None = None, **_unused):
        self._require_cuda()
        if attention_mask is not None and not bool(attention_mask.all()):
            raise NotImplementedError("padding masks are not on the reference's benchmarked path (src/benchmarking/data.py:8-21)")
        if token_type_ids is not None and bool(token_type_ids.any()):
            raise NotImplementedError("token_type_ids != 0 are not on the benchmarked path")
        self.flat.sync_shadow()
        ids = input_ids.to(self.device).contiguous()
        B, S = ids.shape
        train = torch.is_grad_enabled() and self.training
        if labels is not None and train:
            if self._grads_were_dropped():
                self.zero_grad()
            self._step_seed += 1
            self._cur_seed = self._step_seed
            loss = _MaskedLMLossFn.apply(self, ids, labels.to(self.device).contiguous(), self.roberta.embeddings.word_embeddings.weight)
            return ModelOutput(loss=loss, logits=None)
        with torch.no_grad():
            x, _ = self._embed(ids, False)
            for i in range(self.L):
                self._wait_bucket(self._fwd_buckets[1 + i])
                x, _ = self._layer_fwd(i, x, B, S, False)
            logits, _ = self._head_logits(x, keep=False)
            out_logits = logits[:, : self.V].reshape(B, S, self.V) if logits.numel() <= (1 << 28) else None
            if labels is None:
                return ModelOutput(logits=out_logits)
            keep = out_logits.clone() if out_logits is not None else None
            loss, _ = K.cross_entropy_(logits, labels.to(self.device).reshape(-1), V=self.V, write_grad=False)
        return ModelOutput(loss=loss, logits=keep)


class _MaskedLMLossFn(torch.autograd.Function):
    """One autograd node for the whole model (see modeling_gpt_neox._CausalLMLossFn)."""

    @staticmethod
    def forward(ctx, model: B200RobertaForMaskedLM, ids, labels, anchor):
        loss, saved = model._train_forward(ids, labels)
        ctx.model, ctx.saved = model, saved
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        model, saved = ctx.model, ctx.saved
        ctx.saved = None
        if saved is None:
            raise RuntimeError("backward through the B200 model a second time is not supported")
        model._train_backward(saved, grad_out)
        model._ensure_grads_attached()
        return None, None, None, None
