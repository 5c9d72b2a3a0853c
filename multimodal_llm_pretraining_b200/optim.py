"""B200Adam / B200AdamW: fused multi-tensor Adam over the flat parameter store (one kernel launch per step).

Drop-in for the optimizer class hand-off of the reference (`BaseModelClass.optimizer`, src/models/__init__.py:116-126;
constructed by HF Trainer as `cls(param_groups, **kwargs)`, HF:trainer.py:1157-1198): same constructor signature and
semantics as torch.optim.Adam (L2-coupled weight decay — what src/models/pythia.py:43-45 selects, DeepSpeed
`adam_w_mode: False`, src/train.py:157-167) and torch.optim.AdamW (decoupled).  `.param_groups` stay live so that LR
schedulers mutate `group["lr"]` as usual; `.state_dict()` / `.load_state_dict()` carry step, exp_avg, exp_avg_sq.

ZeRO-1: pass `shard=(start, end)` (element range of the flat buffer this rank owns): moments are allocated for the
shard only and the kernel touches only chunks inside it.
"""

from __future__ import annotations

import math
from typing import Any

import torch

from . import kernels as K
from .flat import FlatParams

CHUNK = 65536


class B200Adam(torch.optim.Optimizer):
    adamw_mode = False

    def __init__(self, params, lr: float = 1e-3, betas: tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, shard: tuple[int, int] | None = None, zero_grad_in_step: bool = False, **_ignored):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameters")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        flats = set()
        for g in self.param_groups:
            for p in g["params"]:
                info = getattr(p, "_b200_flat", None)
                if info is None:
                    raise TypeError("B200Adam only optimises parameters of B200 modules (flat fp32 store); "
                                    "use torch.optim.Adam for foreign parameters")
                flats.add(id(info[0]))
                self._flat = info[0]
        if len(flats) != 1:
            raise ValueError("B200Adam expects all parameters to come from one B200 module")
        if len(self.param_groups) > 8:
            raise ValueError("at most 8 param groups are supported by the fused kernel")
        self._shard = shard
        self._zero_grad_in_step = zero_grad_in_step
        self._step = 0
        self._built_for = None
        self._m = self._v = None
        self._p32 = None  # sharded fp32 master (engine shard_master=True): the owned slices, packed like the moments

    # ------------------------------------------------------------------ chunk table / state
    def _build(self) -> None:
        f: FlatParams = self._flat
        dev = f.device
        ranges = self._ranges()
        # state offset of each owned range: ranges are packed back to back in the (sharded) moment buffers
        state_off, acc = [], 0
        for lo, hi in ranges:
            state_off.append(acc)
            acc += hi - lo
        starts, lens, grps, soffs = [], [], [], []
        for gi, g in enumerate(self.param_groups):
            for p in g["params"]:
                _, off, n = p._b200_flat
                for (lo, hi), so in zip(ranges, state_off):
                    a, b = max(off, lo), min(off + n, hi)
                    c = a
                    while c < b:
                        ln = min(CHUNK, b - c)
                        starts.append(c), lens.append(ln), grps.append(gi), soffs.append(so + (c - lo))
                        c += ln
        self._chunk_start = torch.tensor(starts, dtype=torch.int64, device=dev)
        self._chunk_len = torch.tensor(lens, dtype=torch.int32, device=dev)
        self._chunk_group = torch.tensor(grps, dtype=torch.int32, device=dev)
        self._chunk_state = torch.tensor(soffs, dtype=torch.int64, device=dev)
        self._state_base = 0
        if self._m is None or self._m.numel() != acc or self._m.device != dev:
            m_old, v_old = self._m, self._v
            self._m = torch.zeros(acc, dtype=torch.float32, device=dev)
            self._v = torch.zeros(acc, dtype=torch.float32, device=dev)
            if m_old is not None and m_old.numel() == acc:
                self._m.copy_(m_old), self._v.copy_(v_old)
        self._built_for = (dev, self._shadow_ptr(), tuple(ranges))

    def _ranges(self) -> list[tuple[int, int]]:
        if self._shard is None:
            return [(0, self._flat.numel)]
        if isinstance(self._shard, tuple) and len(self._shard) == 2 and isinstance(self._shard[0], int):
            return [self._shard]
        return [tuple(r) for r in self._shard]

    def set_shard(self, ranges) -> None:
        """ZeRO-1/2: restrict the update (and the moment buffers) to the flat-buffer element ranges this rank owns. Must be
        called before any state exists: re-sharding live moments would silently drop them."""
        ranges = [tuple(r) for r in ranges]
        if self._m is not None and self._step > 0 and ranges != self._ranges():
            raise RuntimeError("B200Adam.set_shard: optimizer state already exists for a different shard; build the TrainEngine "
                               "before loading optimizer state (TrainEngine.load_checkpoint does)")
        self._shard = ranges
        self._built_for = None
        self._m = self._v = None

    def chunk_table(self) -> tuple[torch.Tensor, torch.Tensor]:
        """(chunk_start int64, chunk_len int32) device arrays covering exactly the parameter elements this rank updates."""
        self._ensure_built()
        return self._chunk_start, self._chunk_len

    def _ensure_built(self) -> None:
        f = self._flat
        if self._built_for != (f.device, self._shadow_ptr(), tuple(self._ranges())):
            self._build()

    def _shadow_ptr(self) -> int:
        sh = self._flat.shadow  # None under ZeRO-3: the engine casts the updated fp32 shard into its 16-bit shard itself
        return 0 if sh is None else sh.data_ptr()

    def adopt_master_shard(self) -> torch.Tensor:
        """True ZeRO partition of the fp32 weights: copy the slices this rank owns out of the (still full) flat master into a packed
        buffer laid out like the moments; from now on the update reads and writes THAT (the engine then frees the full master)."""
        f = self._flat
        self._ensure_built()
        ranges = self._ranges()
        self._p32 = torch.empty(sum(hi - lo for lo, hi in ranges), dtype=torch.float32, device=f.device)
        off = 0
        for lo, hi in ranges:
            self._p32[off:off + hi - lo].copy_(f.master[lo:hi])
            off += hi - lo
        return self._p32

    def _require_cuda(self) -> None:
        if self._flat.device.type != "cuda":
            raise RuntimeError("B200Adam needs the module on a CUDA (sm_100a) device; there is no CPU fallback")

    # ------------------------------------------------------------------ torch.optim API
    @torch.no_grad()
    def step(self, closure=None, grad_scale: torch.Tensor | None = None, skip_flag: torch.Tensor | None = None,
             grads: torch.Tensor | None = None, grads_packed: bool = False):
        """grad_scale: device fp32 scalar multiplied into every gradient (clip coefficient, fp16 unscale); skip_flag: device
        int32, non-zero = leave parameters and moments untouched (fp16 overflow); grads / grads_packed: a packed shard
        gradient buffer laid out like the moments (ZeRO-2) instead of the flat gradient buffer."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        f = self._flat
        self._require_cuda()
        self._ensure_built()
        self._step += 1
        t = self._step
        groups = []
        for g in self.param_groups:
            b1, b2 = g["betas"]
            groups.append(dict(lr=float(g["lr"]), beta1=b1, beta2=b2, eps=g["eps"], weight_decay=g["weight_decay"],
                               bias_corr1=1.0 - b1 ** t, bias_corr2=1.0 - b2 ** t, adamw_mode=self.adamw_mode))
        if grad_scale is None:
            grad_scale = f.pending_grad_scale
        f.pending_grad_scale = None
        f.sync_shadow()  # no-op unless the master was edited through torch since the last step
        K.adam_step(f.master if self._p32 is None else self._p32, f.grad if grads is None else grads, self._m, self._v, f.shadow,
                    self._state_base, self._chunk_start, self._chunk_len, self._chunk_group, groups, grad_scale=grad_scale,
                    zero_grad=self._zero_grad_in_step, chunk_state=self._chunk_state, skip_flag=skip_flag, g_packed=grads_packed,
                    p_packed=self._p32 is not None)
        return loss

    def zero_grad(self, set_to_none: bool = False) -> None:
        self._flat.zero_grad()

    # ------------------------------------------------------------------ (de)serialisation
    def state_dict(self) -> dict[str, Any]:
        return {
            "step": self._step,
            "shard": self._shard,
            "exp_avg": None if self._m is None else self._m.detach().cpu(),
            "exp_avg_sq": None if self._v is None else self._v.detach().cpu(),
            "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups],
        }

    def load_state_dict(self, sd: dict[str, Any]) -> None:
        self._step = int(sd["step"])
        for g, saved in zip(self.param_groups, sd["param_groups"]):
            g.update(saved)
        if sd.get("exp_avg") is not None:
            self._ensure_built()
            self._m.copy_(sd["exp_avg"])
            self._v.copy_(sd["exp_avg_sq"])


class B200AdamW(B200Adam):
    adamw_mode = True

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2, **kw):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, **kw)


def fused_optimizer_class(opt_cls: type[torch.optim.Optimizer]) -> type[torch.optim.Optimizer]:
    """The model classes hand out the reference's own optimizer classes (torch.optim.Adam for Pythia, src/models/pythia.py:43-45;
    torch.optim.AdamW for RoBERTa, src/models/roberta.py:32-34). For a B200 module the trainer swaps them for the fused
    equivalents with the same constructor signature and update rule; anything else is used as is."""
    if opt_cls is torch.optim.Adam:
        return B200Adam
    if opt_cls is torch.optim.AdamW:
        return B200AdamW
    return opt_cls


# ---------------------------------------------------------------------------------------------------- LR schedules
def cosine_with_min_lr_lambda(step: int, *, num_warmup_steps: int, num_training_steps: int, num_cycles: float = 0.5,
                              min_lr_rate: float = 0.0) -> float:
    """HF:optimization.py:324-334 (_get_cosine_with_min_lr_schedule_with_warmup_lr_lambda)."""
    if step < num_warmup_steps:
        return float(step) / float(max(1, num_warmup_steps))
    progress = float(step - num_warmup_steps) / float(max(1, num_training_steps - num_warmup_steps))
    factor = 0.5 * (1.0 + math.cos(math.pi * float(num_cycles) * 2.0 * progress))
    factor = factor * (1 - min_lr_rate) + min_lr_rate
    return max(0, factor)


def linear_lambda(step: int, *, num_warmup_steps: int, num_training_steps: int) -> float:
    """HF:optimization.py:101-104 (_get_linear_schedule_with_warmup_lr_lambda)."""
    if step < num_warmup_steps:
        return float(step) / float(max(1, num_warmup_steps))
    return max(0.0, float(num_training_steps - step) / float(max(1, num_training_steps - num_warmup_steps)))


def get_scheduler(name: str, optimizer: torch.optim.Optimizer, num_warmup_steps: int, num_training_steps: int,
                  scheduler_specific_kwargs: dict | None = None) -> torch.optim.lr_scheduler.LambdaLR:
    """The two schedules the in-scope model classes use (src/models/pythia.py:69-78, src/models/roberta.py:44-50)."""
    from functools import partial

    kw = dict(scheduler_specific_kwargs or {})
    name = getattr(name, "value", name)
    if name == "cosine_with_min_lr":
        fn = partial(cosine_with_min_lr_lambda, num_warmup_steps=num_warmup_steps, num_training_steps=num_training_steps,
                     num_cycles=kw.get("num_cycles", 0.5), min_lr_rate=kw.get("min_lr_rate", 0.0))
    elif name == "linear":
        fn = partial(linear_lambda, num_warmup_steps=num_warmup_steps, num_training_steps=num_training_steps)
    else:
        raise NotImplementedError(f"scheduler {name!r} is not used by the in-scope model classes")
    return torch.optim.lr_scheduler.LambdaLR(optimizer, fn)
