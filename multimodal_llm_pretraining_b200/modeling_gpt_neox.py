"""B200-native drop-in for transformers.GPTNeoXForCausalLM (the module `src/models/pythia.py:15-22` builds).

Same constructor input (a GPTNeoXConfig-like object), same parameter names / shapes / state_dict keys
(HF:models/gpt_neox/modeling_gpt_neox.py:188-476), same call convention `model(input_ids=..., labels=...)["loss"]`
(src/benchmarking/data.py:17-21, src/benchmarking/flops.py:34), but every FLOP runs in libb200pt (hand-written
sm_100a kernels behind the C ABI): dual-output LayerNorm, tcgen05 GEMMs with fused bias/GELU/residual epilogues,
rotary, tcgen05 flash attention, one-pass cross entropy. Forward and backward are hand-scheduled (one autograd node for
the whole model); parameter gradients are accumulated by the wgrad GEMMs straight into the flat fp32 grad buffer.

Numerics: fp32 master parameters, bf16 compute copies and activations, fp32 accumulation/statistics — the reference's
"bf16 mixed precision" regime.
"""

from __future__ import annotations

import math
import os
from types import SimpleNamespace

import torch
from torch import nn

from . import kernels as K
from .flat import FlatParams

BF16 = torch.bfloat16
# bias gradient of dense_h_to_4h from the dGELU dgrad epilogue (b200_gemm_args.colsum_out) instead of a column-sum pass over dh1.
# Built, parity-tested (tests/test_kernels_gpu.py::test_gemm_epilogues) and measured on the Pythia-1b step: 183.8 / 184.2 k tokens/s
# fused against 185.0 / 184.7 k separate (alternating, one box): the 40 shuffles + 2 vector reductions per 8 columns cost the
# epilogue more than the 90 us pass they replace (in-step GEMM rate 1302 -> 1277 TFLOP/s). Off unless B200_FUSED_BIAS_GRAD=1.
FUSED_BIAS_GRAD = bool(os.environ.get("B200_FUSED_BIAS_GRAD"))


class ModelOutput(dict):
    """dict with attribute access: supports outputs["loss"], outputs.loss and outputs.get("loss") like HF ModelOutput."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __getitem__(self, k):
        if isinstance(k, int):  # HF Trainer does outputs[0] when labels are absent from a dict-less output
            return [v for v in self.values() if v is not None][k]
        return super().__getitem__(k)


class _Params(nn.Module):
    """Leaf container so that parameter names match HF's module tree (e.g. `attention.query_key_value.weight`)."""

    def __init__(self, flat: FlatParams, prefix: str, names: tuple[str, ...]):
        super().__init__()
        for n in names:
            self.register_parameter(n, flat.make_parameter(f"{prefix}.{n}"))


def neox_param_shapes(cfg) -> list[tuple[str, tuple[int, ...]]]:
    h, V, L, I = cfg.hidden_size, cfg.vocab_size, cfg.num_hidden_layers, cfg.intermediate_size
    shapes: list[tuple[str, tuple[int, ...]]] = [("gpt_neox.embed_in.weight", (V, h))]
    for i in range(L):
        p = f"gpt_neox.layers.{i}"
        shapes += [
            (f"{p}.input_layernorm.weight", (h,)), (f"{p}.input_layernorm.bias", (h,)),
            (f"{p}.post_attention_layernorm.weight", (h,)), (f"{p}.post_attention_layernorm.bias", (h,)),
            (f"{p}.attention.query_key_value.weight", (3 * h, h)), (f"{p}.attention.query_key_value.bias", (3 * h,)),
            (f"{p}.attention.dense.weight", (h, h)), (f"{p}.attention.dense.bias", (h,)),
            (f"{p}.mlp.dense_h_to_4h.weight", (I, h)), (f"{p}.mlp.dense_h_to_4h.bias", (I,)),
            (f"{p}.mlp.dense_4h_to_h.weight", (h, I)), (f"{p}.mlp.dense_4h_to_h.bias", (h,)),
        ]
    shapes += [("gpt_neox.final_layer_norm.weight", (h,)), ("gpt_neox.final_layer_norm.bias", (h,)),
               ("embed_out.weight", (V, h))]
    return shapes


class _FlatModule(nn.Module):
    """nn.Module whose parameters are views into a FlatParams store; keeps the views intact across .to()/.cuda()."""

    flat: FlatParams

    def _named_flat_params(self) -> dict[str, nn.Parameter]:
        return dict(self.named_parameters())

    def _apply(self, fn, recurse=True):  # noqa: D401 - nn.Module hook
        self.flat.apply(fn)
        self.flat.rebind(self._named_flat_params())
        if hasattr(self, "_rope_cache"):
            self._rope_cache = {}
        return self

    # ---- precision: bf16 (default) or fp16 + loss scaling (src/models/pythia.py:33-41: fp16 for every Pythia but 1b)
    loss_scale: torch.Tensor | None = None  # device fp32 scalar set by the engine's LossScaler in fp16 runs

    @property
    def compute_dtype(self) -> torch.dtype:
        return self.flat.compute_dtype

    def set_compute_dtype(self, dtype: torch.dtype) -> None:
        """torch.bfloat16 or torch.float16: selects libb200pt.so / libb200pt_fp16.so for every kernel of this module. fp16
        needs a loss scale (engine.LossScaler sets `loss_scale`; the cross entropy folds it into dlogits, the optimizer's
        clip coefficient divides it out again)."""
        self.flat.set_compute_dtype(dtype)

    def zero_grad(self, set_to_none: bool = False) -> None:  # grads are persistent views; "None" means zero here
        self.flat.zero_grad()
        self._ensure_grads_attached()

    def state_dict(self, *args, **kwargs):
        # ZeRO-1 replicates only the bf16 compute copy between steps; bring every rank's fp32 master up to date first
        # (a collective: like DeepSpeed's consolidated state_dict, all ranks must call it)
        if self.flat.master is None:  # sharded master: fp32 values all-gathered from their owners into a fresh full copy
            from collections import OrderedDict

            full = self.flat.materialize_master()
            prefix = kwargs.get("prefix", args[1] if len(args) > 1 else "")
            return OrderedDict((prefix + n, self.flat.view(full, n)) for n, _ in self.named_parameters())
        self.flat.consolidate()
        return super().state_dict(*args, **kwargs)

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        if self.flat.master is None:  # sharded master: scatter a full fp32 copy back to the owners (collective)
            full = self.flat.materialize_master()
            own = [n for n, _ in self.named_parameters()]
            missing = [n for n in own if n not in state_dict]
            if strict and missing:
                raise KeyError(f"missing keys: {missing[:4]}")
            with torch.no_grad():
                for n in own:
                    if n in state_dict:
                        self.flat.view(full, n).copy_(state_dict[n].to(torch.float32))
            self.flat.master_loader(full)
            return None
        return super().load_state_dict(state_dict, strict=strict, assign=assign)

    def _ensure_grads_attached(self) -> None:
        if self.flat.grad is None or self.flat.master is None:
            return  # ZeRO-2 (gradients live in transient bucket buffers / the shard accumulator) or a sharded master (16-bit parameters)
        for name, p in self.named_parameters():
            if p.grad is None:
                p.grad = self.flat.view(self.flat.grad, name)

    def _grads_were_dropped(self) -> bool:
        """A foreign zero_grad(set_to_none=True) (torch optimizers, nn.Module default) drops the views: treat as zero."""
        if self.flat.grad is None or self.flat.master is None:
            return False
        p = next(self.parameters())
        return p.grad is None

    # engine hooks: grad_ready_hook(start, end) fires in backward as each bucket's gradients complete; param_wait_hook(start,
    # end) is called in forward right before the parameters of a bucket are first read (ZeRO: the bucket's all-gather, which
    # runs on a side stream, must have landed)
    grad_ready_hook = None
    param_wait_hook = None

    def _require_cuda(self) -> None:
        """The kernels exist for sm_100a only; fail before any of them is reached with a message that says what to do."""
        if self.flat.device.type != "cuda":
            raise RuntimeError(f"{type(self).__name__} runs only on a CUDA (sm_100a) device: there is no CPU fallback. "
                               "Move the module with .cuda() first.")

    def _wait_bucket(self, rng: tuple[int, int]) -> None:
        if self.param_wait_hook is not None:
            self.param_wait_hook(*rng)

    # ZeRO-3 only: called in BACKWARD right before the weights of a bucket are read again (dgrad GEMMs, recomputation); None otherwise
    param_bwd_hook = None
    supports_weight_sharding = False  # True on modules whose backward announces its weight reads through `_need_bucket_bwd`

    def _need_bucket_bwd(self, rng: tuple[int, int]) -> None:
        if self.param_bwd_hook is not None:
            self.param_bwd_hook(*rng)

    def clip_grad_norm_(self, max_norm: float) -> torch.Tensor:
        """Fused replacement for torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)
        (src/benchmarking/utils.py:66-70): one sum-of-squares pass over the flat grad buffer; the clip coefficient stays
        on the device and is folded into the next B200Adam.step(). Returns the total norm (device scalar)."""
        f = self.flat
        if f.grad is None:
            raise RuntimeError("gradients are sharded by the TrainEngine (ZeRO-2): clipping happens in manual_optimization_step")
        sumsq = torch.zeros((), dtype=torch.float32, device=f.grad.device)
        K.sumsq_(f.grad, sumsq)
        norm, coef = K.clip_coef(sumsq, max_norm)
        f.pending_grad_scale = coef
        return norm


class _HeadGrad:
    __slots__ = ("xf", "dlogits", "x_last", "mean", "rstd")


class B200GPTNeoXForCausalLM(_FlatModule):
    supports_gradient_checkpointing = True
    supports_weight_sharding = True
    main_input_name = "input_ids"

    def __init__(self, config):
        super().__init__()
        self.config = config
        cfg = config
        assert getattr(cfg, "use_parallel_residual", True), "only the parallel-residual GPT-NeoX block (all Pythia models) is built"
        assert getattr(cfg, "hidden_act", "gelu") == "gelu", "only exact-erf GELU is built"
        assert getattr(cfg, "attention_bias", True), "attention_bias=False is not built"
        assert not getattr(cfg, "tie_word_embeddings", False), "GPT-NeoX/Pythia has untied embed_out"
        self.h = cfg.hidden_size
        self.nh = cfg.num_attention_heads
        self.hd = self.h // self.nh
        rope = getattr(cfg, "rope_parameters", None) or {}
        pct = rope.get("partial_rotary_factor", getattr(cfg, "rotary_pct", 0.25))
        self.rope_base = float(rope.get("rope_theta", getattr(cfg, "rotary_emb_base", 10000)))
        self.rot = int(self.hd * pct)
        self.eps = cfg.layer_norm_eps
        self.L = cfg.num_hidden_layers
        self.V = cfg.vocab_size
        self.inter = cfg.intermediate_size
        if self.hd not in (64, 80, 128, 256):
            raise NotImplementedError(f"head_dim {self.hd} has no tcgen05 attention kernel yet (built: 64, 80, 128, 256)")

        self.flat = FlatParams(neox_param_shapes(cfg))
        f = self.flat
        body = nn.Module()
        body.embed_in = _Params(f, "gpt_neox.embed_in", ("weight",))
        layers = []
        for i in range(self.L):
            p = f"gpt_neox.layers.{i}"
            lyr = nn.Module()
            lyr.input_layernorm = _Params(f, f"{p}.input_layernorm", ("weight", "bias"))
            lyr.post_attention_layernorm = _Params(f, f"{p}.post_attention_layernorm", ("weight", "bias"))
            att = nn.Module()
            att.query_key_value = _Params(f, f"{p}.attention.query_key_value", ("weight", "bias"))
            att.dense = _Params(f, f"{p}.attention.dense", ("weight", "bias"))
            lyr.attention = att
            mlp = nn.Module()
            mlp.dense_h_to_4h = _Params(f, f"{p}.mlp.dense_h_to_4h", ("weight", "bias"))
            mlp.dense_4h_to_h = _Params(f, f"{p}.mlp.dense_4h_to_h", ("weight", "bias"))
            lyr.mlp = mlp
            layers.append(lyr)
        body.layers = nn.ModuleList(layers)
        body.final_layer_norm = _Params(f, "gpt_neox.final_layer_norm", ("weight", "bias"))
        self.gpt_neox = body
        self.embed_out = _Params(f, "embed_out", ("weight",))
        self.gradient_checkpointing = False
        self.grad_ready_hook = None  # callable(start, end) on flat-grad element ranges, fired in backward order
        self.param_wait_hook = None
        self._layer_ranges = [self._layer_range(i) for i in range(self.L)]
        self._rope_cache: dict[tuple, tuple[torch.Tensor, torch.Tensor]] = {}
        self.reset_parameters()

    # ------------------------------------------------------------------ init / HF surface
    @torch.no_grad()
    def reset_parameters(self, generator: torch.Generator | None = None) -> None:
        """HF _init_weights for GPT-NeoX: Linear/Embedding N(0, initializer_range), biases 0, LayerNorm (1, 0)."""
        std = getattr(self.config, "initializer_range", 0.02)
        for name, p in self.named_parameters():
            if name.endswith("layernorm.weight") or name.endswith("layer_norm.weight"):
                p.fill_(1.0)
            elif name.endswith(".bias"):
                p.zero_()
            else:
                p.normal_(0.0, std, generator=generator)

    def gradient_checkpointing_enable(self, gradient_checkpointing_kwargs=None) -> None:
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
        return self.gpt_neox.embed_in

    def get_output_embeddings(self):
        return self.embed_out

    def num_parameters(self, only_trainable: bool = False) -> int:
        return sum(p._b200_flat[2] for p in self.parameters())  # the store's record: valid while the tensors are sharded away too

    def load_hf_state_dict(self, sd: dict[str, torch.Tensor]) -> None:
        own = self.state_dict()
        missing = [k for k in own if k not in sd]
        if missing:
            raise KeyError(f"missing keys: {missing[:4]}…")
        with torch.no_grad():
            for k, v in own.items():
                v.copy_(sd[k].to(v.dtype))

    # ------------------------------------------------------------------ helpers
    def _w(self, name: str) -> torch.Tensor:  # bf16 compute copy (ZeRO-3: the bucket buffer the engine gathered it into)
        return self.flat.wview(name)

    def _p(self, name: str) -> torch.Tensor:  # fp32 values the kernels read directly (LayerNorm affine, biases)
        return self.flat.pview(name)

    def _g(self, name: str) -> torch.Tensor:  # fp32 gradient accumulator
        return self.flat.gview(name)

    def _rope_tables(self, S: int, device) -> tuple[torch.Tensor, torch.Tensor]:
        key = (S, device)
        if key not in self._rope_cache:
            # HF GPTNeoXRotaryEmbedding (modeling_gpt_neox.py:52-116): inv_freq = base^(-2i/rot); angles in fp32
            inv = 1.0 / (self.rope_base ** (torch.arange(0, self.rot, 2, dtype=torch.int64).float() / self.rot))
            ang = torch.arange(S, dtype=torch.float32)[:, None] * inv[None, :]
            self._rope_cache = {key: (ang.cos().contiguous().to(device), ang.sin().contiguous().to(device))}
        return self._rope_cache[key]

    # ------------------------------------------------------------------ one transformer layer
    def _layer_fwd(self, i: int, x: torch.Tensor, B: int, S: int, keep: bool):
        """x bf16 [T,h] -> y bf16 [T,h]; returns (y, saved) where saved holds what backward needs (None if not keep)."""
        p = f"gpt_neox.layers.{i}"
        h, nh, hd = self.h, self.nh, self.hd
        a1, a2, mean, rstd = K.layernorm_fwd(x, self._p(f"{p}.input_layernorm.weight"), self._p(f"{p}.input_layernorm.bias"), self.eps,
                                             self._p(f"{p}.post_attention_layernorm.weight"), self._p(f"{p}.post_attention_layernorm.bias"))
        qkv = K.gemm(a1, self._w(f"{p}.attention.query_key_value.weight"), bias=self._p(f"{p}.attention.query_key_value.bias"))
        cos, sin = self._rope_tables(S, x.device)
        K.rope_qk_inplace(qkv, cos, sin, B, S, nh, hd, self.rot)
        qkv5 = qkv.view(B, S, nh, 3, hd)
        o, lse = K.attention_fwd(qkv5[:, :, :, 0], qkv5[:, :, :, 1], qkv5[:, :, :, 2], causal=True, scale=hd ** -0.5)
        o2 = o.view(B * S, h)
        att = K.gemm(o2, self._w(f"{p}.attention.dense.weight"), bias=self._p(f"{p}.attention.dense.bias"), residual=x)
        h1 = torch.empty(B * S, self.inter, dtype=x.dtype, device=x.device) if keep else None
        g = K.gemm(a2, self._w(f"{p}.mlp.dense_h_to_4h.weight"), bias=self._p(f"{p}.mlp.dense_h_to_4h.bias"), gelu=True, aux_out=h1)
        y = K.gemm(g, self._w(f"{p}.mlp.dense_4h_to_h.weight"), bias=self._p(f"{p}.mlp.dense_4h_to_h.bias"), residual=att)
        saved = (x, mean, rstd, a1, a2, qkv, o, lse, h1, g) if keep else None
        return y, saved

    def _layer_bwd(self, i: int, saved, dy: torch.Tensor, B: int, S: int) -> torch.Tensor:
        p = f"gpt_neox.layers.{i}"
        x, mean, rstd, a1, a2, qkv, o, lse, h1, g = saved
        h, nh, hd = self.h, self.nh, self.hd
        T = B * S
        # --- MLP branch: y = W2 gelu(W1 a2 + b1) + b2
        K.gemm(dy, g, a_mn=True, b_mn=True, out=self._g(f"{p}.mlp.dense_4h_to_h.weight"), accumulate=True)
        # both output biases of the parallel-residual block see the same upstream gradient: one column-sum pass, two accumulators
        K.colsum_(dy, self._g(f"{p}.mlp.dense_4h_to_h.bias"), out2=self._g(f"{p}.attention.dense.bias"))
        # dgrad through W2 with gelu'(h1) and the bias gradient of dense_h_to_4h (column sums of dh1) in the same epilogue
        if FUSED_BIAS_GRAD:
            dh1 = K.gemm(dy, self._w(f"{p}.mlp.dense_4h_to_h.weight"), b_mn=True, dgelu_in=h1, colsum_out=self._g(f"{p}.mlp.dense_h_to_4h.bias"))
        else:
            dh1 = K.gemm(dy, self._w(f"{p}.mlp.dense_4h_to_h.weight"), b_mn=True, dgelu_in=h1)
            K.colsum_(dh1, self._g(f"{p}.mlp.dense_h_to_4h.bias"))
        K.gemm(dh1, a2, a_mn=True, b_mn=True, out=self._g(f"{p}.mlp.dense_h_to_4h.weight"), accumulate=True)
        da2 = K.gemm(dh1, self._w(f"{p}.mlp.dense_h_to_4h.weight"), b_mn=True)
        del dh1
        # --- attention branch
        o2 = o.view(T, h)
        K.gemm(dy, o2, a_mn=True, b_mn=True, out=self._g(f"{p}.attention.dense.weight"), accumulate=True)
        d_o = K.gemm(dy, self._w(f"{p}.attention.dense.weight"), b_mn=True)
        dqkv = torch.empty_like(qkv)
        q5, d5 = qkv.view(B, S, nh, 3, hd), dqkv.view(B, S, nh, 3, hd)
        K.attention_bwd(q5[:, :, :, 0], q5[:, :, :, 1], q5[:, :, :, 2], o, lse, d_o.view(B, S, nh, hd),
                        d5[:, :, :, 0], d5[:, :, :, 1], d5[:, :, :, 2], causal=True, scale=hd ** -0.5)
        cos, sin = self._rope_tables(S, x.device)
        K.rope_qk_inplace(dqkv, cos, sin, B, S, nh, hd, self.rot, inverse=True)
        K.gemm(dqkv, a1, a_mn=True, b_mn=True, out=self._g(f"{p}.attention.query_key_value.weight"), accumulate=True)
        K.colsum_(dqkv, self._g(f"{p}.attention.query_key_value.bias"))
        da1 = K.gemm(dqkv, self._w(f"{p}.attention.query_key_value.weight"), b_mn=True)
        # --- both LayerNorms share x: one fused backward, residual gradient added in the same pass
        dx = K.layernorm_bwd(x, mean, rstd, self._p(f"{p}.input_layernorm.weight"), da1,
                             self._g(f"{p}.input_layernorm.weight"), self._g(f"{p}.input_layernorm.bias"),
                             self._p(f"{p}.post_attention_layernorm.weight"), da2,
                             self._g(f"{p}.post_attention_layernorm.weight"), self._g(f"{p}.post_attention_layernorm.bias"),
                             dres=dy)
        return dx

    def _layer_range(self, i: int) -> tuple[int, int]:
        p = f"gpt_neox.layers.{i}"
        return self.flat.range_of([n for n in self.flat.names if n.startswith(p + ".")])

    def _head_range(self) -> tuple[int, int]:
        return self.flat.range_of(["gpt_neox.final_layer_norm.weight", "gpt_neox.final_layer_norm.bias", "embed_out.weight"])

    def comm_buckets(self) -> list[tuple[int, int]]:
        """Flat-grad element ranges in the order backward completes them (head, layers L-1..0, input embedding); these
        are the DDP / ZeRO-1 communication buckets `grad_ready_hook` is fired with."""
        return [self._head_range()] + [self._layer_range(i) for i in reversed(range(self.L))] + \
               [self.flat.range_of(["gpt_neox.embed_in.weight"])]

    # ------------------------------------------------------------------ whole-model forward / backward
    def _forward_hidden(self, ids: torch.Tensor, keep: bool):
        B, S = ids.shape
        self._wait_bucket(self.flat.range_of(["gpt_neox.embed_in.weight"]))
        x = K.embedding_fwd(ids.reshape(-1), self._w("gpt_neox.embed_in.weight"))
        saved_layers = []
        for i in range(self.L):
            self._wait_bucket(self._layer_ranges[i])
            if keep and self.gradient_checkpointing:
                saved_layers.append(x)  # recompute the layer in backward
                x, _ = self._layer_fwd(i, x, B, S, keep=False)
            else:
                x, sv = self._layer_fwd(i, x, B, S, keep=keep)
                saved_layers.append(sv)
        self._wait_bucket(self._head_range())
        xf, _, mean, rstd = K.layernorm_fwd(x, self._p("gpt_neox.final_layer_norm.weight"), self._p("gpt_neox.final_layer_norm.bias"), self.eps)
        return x, xf, mean, rstd, saved_layers

    def _train_forward(self, ids: torch.Tensor, targets: torch.Tensor):
        """ids, targets int64 [B,S] (already shifted). Returns (loss, ctx for backward)."""
        B, S = ids.shape
        x_last, xf, mean, rstd, saved_layers = self._forward_hidden(ids, keep=True)
        logits = K.gemm(xf, self._w("embed_out.weight"))  # [T, V] bf16 — overwritten in place by dlogits
        loss, _ = K.cross_entropy_(logits, targets.reshape(-1), V=self.V, write_grad=True, grad_scale=self.loss_scale)
        ctx = SimpleNamespace(ids=ids, B=B, S=S, saved_layers=saved_layers, x_last=x_last, xf=xf, mean=mean, rstd=rstd, dlogits=logits)
        return loss, ctx

    def _train_backward(self, ctx, grad_out: torch.Tensor) -> None:
        B, S = ctx.B, ctx.S
        alpha = grad_out.reshape(1).to(torch.float32).contiguous()
        hook = self.grad_ready_hook
        self._need_bucket_bwd(self._head_range())
        # LM head: dWout += alpha * dlogits^T xf ; dxf = alpha * dlogits Wout
        K.gemm(ctx.dlogits, ctx.xf, a_mn=True, b_mn=True, out=self._g("embed_out.weight"), accumulate=True, alpha=alpha)
        dxf = K.gemm(ctx.dlogits, self._w("embed_out.weight"), b_mn=True, alpha=alpha)
        ctx.dlogits = None
        dx = K.layernorm_bwd(ctx.x_last, ctx.mean, ctx.rstd, self._p("gpt_neox.final_layer_norm.weight"), dxf,
                             self._g("gpt_neox.final_layer_norm.weight"), self._g("gpt_neox.final_layer_norm.bias"))
        if hook:
            hook(*self.flat.range_of(["gpt_neox.final_layer_norm.weight", "gpt_neox.final_layer_norm.bias", "embed_out.weight"]))
        for i in reversed(range(self.L)):
            sv = ctx.saved_layers[i]
            self._need_bucket_bwd(self._layer_ranges[i])
            if isinstance(sv, torch.Tensor):  # checkpointed: recompute this layer's activations
                _, sv = self._layer_fwd(i, sv, B, S, keep=True)
            dx = self._layer_bwd(i, sv, dx, B, S)
            ctx.saved_layers[i] = None
            if hook:
                hook(*self._layer_range(i))
        K.embedding_bwd(ctx.ids.reshape(-1), dx, self._g("gpt_neox.embed_in.weight"))
        if hook:
            hook(*self.flat.range_of(["gpt_neox.embed_in.weight"]))

    def forward(self, input_ids: torch.Tensor, labels: torch.Tensor | None = None, attention_mask: torch.Tensor | None = None, **_unused):
        self._require_cuda()
        if attention_mask is not None and not bool(attention_mask.all()):
            raise NotImplementedError("padding masks are not on the reference's benchmarked path (src/benchmarking/data.py:8-21)")
        self.flat.sync_shadow()
        input_ids = input_ids.to(self.device)
        if labels is None:
            with torch.no_grad():
                B, S = input_ids.shape
                _, xf, _, _, _ = self._forward_hidden(input_ids.contiguous(), keep=False)
                logits = K.gemm(xf, self._w("embed_out.weight")).view(B, S, self.V)
            return ModelOutput(logits=logits)
        labels = labels.to(self.device)
        # HF ForCausalLMLoss shifts inside the model (loss_utils.py:45-67): position t predicts labels[t+1]; the last
        # position's logits are never used and, under causal attention, feed no other position -> drop it up front
        # (SURVEY.md App. B.3): 2049 tokens in, 2048 computed.
        ids = input_ids[:, :-1].contiguous()
        targets = labels[:, 1:].contiguous()
        if torch.is_grad_enabled() and self.training:
            if self._grads_were_dropped():
                self.zero_grad()
            loss = _CausalLMLossFn.apply(self, ids, targets, self.gpt_neox.embed_in.weight)
            return ModelOutput(loss=loss, logits=None)
        with torch.no_grad():
            B, S = ids.shape
            _, xf, _, _, _ = self._forward_hidden(ids, keep=False)
            logits = K.gemm(xf, self._w("embed_out.weight"))
            keep_logits = logits.clone() if logits.numel() <= (1 << 28) else None
            loss, _ = K.cross_entropy_(logits, targets.reshape(-1), V=self.V, write_grad=False)
        return ModelOutput(loss=loss, logits=None if keep_logits is None else keep_logits.view(B, S, self.V))


class _CausalLMLossFn(torch.autograd.Function):
    """One autograd node for the whole model: forward = hand-scheduled kernel sequence, backward likewise; parameter
    gradients are written by the kernels into the flat grad buffer (the `anchor` parameter only ties the node into the
    graph so that loss.backward() reaches it)."""

    @staticmethod
    def forward(ctx, model: B200GPTNeoXForCausalLM, ids, targets, anchor):
        loss, saved = model._train_forward(ids, targets)
        ctx.model = model
        ctx.saved = saved
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
