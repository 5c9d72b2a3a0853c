"""Flat parameter storage for the B200 modules.

HBM layout (one contiguous buffer each, same offsets in all of them):
    master : fp32   — the nn.Parameters are views into it (HF-compatible state_dict, torch optimizers work on it)
    grad   : fp32   — the .grad of every parameter is a view into it; wgrad GEMMs accumulate straight into it
    shadow : bf16   — what the tensor-core kernels read; rewritten by the fused Adam in the same pass that updates master
             (fp16 after set_compute_dtype(torch.float16): the reference's precision for every Pythia but 1b and RoBERTa)
Every parameter starts at a multiple of 64 elements (256 B fp32 / 128 B bf16) so TMA bases and vector accesses are aligned.
A contiguous layout makes the optimizer one launch, the gradient norm one launch, and DDP / ZeRO-1 collectives plain
slices of one buffer (bucket = a contiguous range, shard = a contiguous range).
"""

from __future__ import annotations

import math

import torch
from torch import nn

ALIGN = 64


class FlatParams:
    def __init__(self, shapes: list[tuple], device="cpu", compute_dtype: torch.dtype = torch.bfloat16):
        """shapes: (name, shape) or (name, shape, alloc_shape): alloc_shape >= shape reserves zero padding behind the
        parameter (e.g. RoBERTa's 50265-row vocabulary padded to 50304 rows so that the tied decoder GEMM, its wgrad and
        the cross entropy run on 16-byte-aligned rows); the nn.Parameter and state_dict see `shape` only."""
        self.names = [e[0] for e in shapes]
        self.shapes = {e[0]: tuple(e[1]) for e in shapes}
        self.alloc_shapes = {e[0]: tuple(e[2]) if len(e) > 2 else tuple(e[1]) for e in shapes}
        self.offsets: dict[str, int] = {}
        off = 0
        for n in self.names:
            self.offsets[n] = off
            off += (math.prod(self.alloc_shapes[n]) + ALIGN - 1) // ALIGN * ALIGN
        # total padded to a multiple of 8 * ALIGN so it splits evenly into up to 8 aligned ZeRO shards
        self.numel = (off + 8 * ALIGN - 1) // (8 * ALIGN) * (8 * ALIGN)
        self.master = torch.zeros(self.numel, dtype=torch.float32, device=device)
        self.grad = torch.zeros(self.numel, dtype=torch.float32, device=device)
        self.compute_dtype = compute_dtype
        self.shadow = torch.zeros(self.numel, dtype=compute_dtype, device=device)
        self.shadow_version = -1
        self.params: list[nn.Parameter] = []
        self.pending_grad_scale: torch.Tensor | None = None  # clip coefficient folded into the next fused Adam step
        # ZeRO-1 (engine.py): between optimizer steps only the owner's fp32 master slice and the replicated bf16 shadow
        # are current; `master_consolidator` (a collective) refreshes the rest of the master on demand
        self.master_stale = False
        self.master_consolidator = None
        # ZeRO-2 (engine.py): there is no full gradient buffer (`grad` is None); `grad_router(name)` returns the transient
        # bucket buffer (and the offset inside it) the gradient of `name` is accumulated into during the current backward,
        # `grad_zero_fn()` clears the rank's shard accumulator
        self.grad_router = None
        self.grad_zero_fn = None
        # Sharded fp32 master (engine.py, shard_master=True): `master` is None — each rank keeps the fp32 values of the slices it owns
        # inside its optimizer (4 B/param/W) — the nn.Parameters become views of the replicated 16-bit copy (as under DeepSpeed, where
        # the module is bf16 and the fp32 weights live in the ZeRO partition), and the few fp32 values the kernels read directly
        # (every 1-D parameter: biases, LayerNorm affine) live in `small`, a compact replicated fp32 buffer.
        # `master_materializer()` (collective) returns a full fp32 copy on demand, `master_loader(full)` scatters one back.
        self.small: torch.Tensor | None = None
        self.small_offsets: dict[str, int] = {}
        self.master_materializer = None
        self.master_loader = None
        # ZeRO-3 (engine.py, strategy "zero3"): the 16-bit compute copy is sharded as well (`shadow` is None; each rank keeps the slices
        # it owns, 2 B/param/W); `weight_router(name)` returns the transient bucket buffer (and the offset inside it) into which the
        # engine has all-gathered the bucket of `name` for the layer that is about to run.
        self.weight_router = None

    @property
    def device(self) -> torch.device:
        for t in (self.master, self.shadow, self.small):
            if t is not None:
                return t.device
        raise RuntimeError("empty parameter store")

    # ---- 16-bit compute-copy views (through the router when the engine shards the weights)
    def _wbuf(self, name: str) -> tuple[torch.Tensor, int]:
        if self.weight_router is not None:
            return self.weight_router(name)
        return self.shadow, self.offsets[name]

    def wview(self, name: str) -> torch.Tensor:
        buf, o = self._wbuf(name)
        s = self.shapes[name]
        return buf[o:o + math.prod(s)].view(s)

    def wview_alloc(self, name: str) -> torch.Tensor:
        buf, o = self._wbuf(name)
        s = self.alloc_shapes[name]
        return buf[o:o + math.prod(s)].view(s)

    def wview_span(self, first: str, shape: tuple[int, ...]) -> torch.Tensor:
        buf, o = self._wbuf(first)
        return buf[o:o + math.prod(shape)].view(shape)

    def drop_shadow(self, params: dict[str, nn.Parameter]) -> None:
        """Free the full 16-bit copy (ZeRO-3; the master must already be sharded). The nn.Parameters become empty placeholders that
        keep their identity, names and `_b200_flat` record — like DeepSpeed stage 3, where `param.data` is an empty tensor between
        uses and `ds_shape` / `ds_numel` carry the real geometry."""
        if self.master is not None:
            raise RuntimeError("drop_shadow: shard the fp32 master first (the parameters would still be views of it)")
        dtype, dev = self.shadow.dtype, self.shadow.device
        for name, p_ in params.items():
            p_.ds_shape, p_.ds_numel = self.shapes[name], math.prod(self.shapes[name])  # type: ignore[attr-defined]
            p_.grad = None
            p_.data = torch.empty(0, dtype=dtype, device=dev)
        self.shadow = None

    # ---- fp32 views of the parameters the kernels read in fp32 (1-D: biases, LayerNorm affine)
    def pview(self, name: str) -> torch.Tensor:
        if self.master is not None:
            return self.view(self.master, name)
        o, s = self.small_offsets[name], self.shapes[name]
        return self.small[o:o + math.prod(s)].view(s)

    def pview_alloc(self, name: str) -> torch.Tensor:
        if self.master is not None:
            return self.view_alloc(self.master, name)
        o, s = self.small_offsets[name], self.alloc_shapes[name]
        return self.small[o:o + math.prod(s)].view(s)

    def pview_span(self, first: str, shape: tuple[int, ...]) -> torch.Tensor:
        if self.master is not None:
            return self.view_span(self.master, first, shape)
        o = self.small_offsets[first]
        return self.small[o:o + math.prod(shape)].view(shape)

    def small_names(self) -> list[str]:
        return [n for n in self.names if len(self.shapes[n]) < 2]

    def drop_master(self, params: dict[str, nn.Parameter]) -> None:
        """Free the full fp32 master: keep the 1-D parameters in `small` (same order, alloc sizes rounded up to ALIGN so that spans
        of neighbouring 1-D parameters stay contiguous and 16-byte aligned), re-point the nn.Parameters at the 16-bit copy."""
        off = 0
        for n in self.small_names():
            self.small_offsets[n] = off
            off += (math.prod(self.alloc_shapes[n]) + ALIGN - 1) // ALIGN * ALIGN
        self.small = torch.zeros(max(off, 1), dtype=torch.float32, device=self.master.device)
        for n in self.small_names():
            k = math.prod(self.alloc_shapes[n])
            self.small[self.small_offsets[n]:self.small_offsets[n] + k].copy_(self.master[self.offsets[n]:self.offsets[n] + k])
        self.master = None
        for name, p_ in params.items():
            p_.grad = None  # fp32 gradient views cannot hang off 16-bit parameters; the kernels reach them through gview()
            p_.data = self.view(self.shadow, name)
        self.master_stale = False

    def materialize_master(self) -> torch.Tensor:
        """Full fp32 parameter vector (flat layout): the master itself, or — sharded — a fresh copy all-gathered from the owners
        (collective: every rank must call it)."""
        if self.master is not None:
            self.consolidate()
            return self.master
        if self.master_materializer is None:
            raise RuntimeError("the fp32 master was dropped but no materializer is installed")
        return self.master_materializer()

    # ---- gradient views (through the router when the engine shards gradients)
    def _gbuf(self, name: str) -> tuple[torch.Tensor, int]:
        if self.grad_router is not None:
            return self.grad_router(name)
        return self.grad, self.offsets[name]

    def gview(self, name: str) -> torch.Tensor:
        buf, o = self._gbuf(name)
        s = self.shapes[name]
        return buf[o:o + math.prod(s)].view(s)

    def gview_alloc(self, name: str) -> torch.Tensor:
        buf, o = self._gbuf(name)
        s = self.alloc_shapes[name]
        return buf[o:o + math.prod(s)].view(s)

    def gview_span(self, first: str, shape: tuple[int, ...]) -> torch.Tensor:
        buf, o = self._gbuf(first)
        return buf[o:o + math.prod(shape)].view(shape)

    # ---- views
    def view(self, buf: torch.Tensor, name: str) -> torch.Tensor:
        o, s = self.offsets[name], self.shapes[name]
        return buf[o:o + math.prod(s)].view(s)

    def view_alloc(self, buf: torch.Tensor, name: str) -> torch.Tensor:
        """View including the zero padding reserved behind the parameter (alloc_shape)."""
        o, s = self.offsets[name], self.alloc_shapes[name]
        return buf[o:o + math.prod(s)].view(s)

    def view_span(self, buf: torch.Tensor, first: str, shape: tuple[int, ...]) -> torch.Tensor:
        """View of `shape` starting at parameter `first` and running over the parameters laid out behind it (e.g. RoBERTa's
        query/key/value weights, contiguous in the store, as one [3h, h] matrix)."""
        o = self.offsets[first]
        return buf[o:o + math.prod(shape)].view(shape)

    def range_of(self, names: list[str]) -> tuple[int, int]:
        """[start, end) element range covering the (contiguous) parameters `names`, including alignment padding."""
        starts = [self.offsets[n] for n in names]
        ends = [self.offsets[n] + (math.prod(self.alloc_shapes[n]) + ALIGN - 1) // ALIGN * ALIGN for n in names]
        return min(starts), max(ends)

    def make_parameter(self, name: str) -> nn.Parameter:
        p = nn.Parameter(self.view(self.master, name), requires_grad=True)
        p.grad = self.view(self.grad, name) if self.grad is not None else None
        p._b200_flat = (self, self.offsets[name], math.prod(self.shapes[name]))  # type: ignore[attr-defined]
        self.params.append(p)
        return p

    def rebind(self, params: dict[str, nn.Parameter]) -> None:
        for name, p in params.items():
            p.data = self.view(self.master, name)
            p.grad = self.view(self.grad, name) if self.grad is not None else None
            p._b200_flat = (self, self.offsets[name], math.prod(self.shapes[name]))  # type: ignore[attr-defined]

    def apply(self, fn) -> None:
        """Move/cast the buffers (used by nn.Module._apply: .to(), .cuda()). The master stays fp32."""
        if self.master is None:
            raise RuntimeError("a module whose fp32 master is sharded by the TrainEngine cannot be moved or cast")
        new_master = fn(self.master)
        if new_master.dtype != torch.float32:
            raise TypeError("B200 modules keep fp32 master parameters; bf16 compute copies are managed internally "
                            "(do not call .half()/.bfloat16() on the module)")
        self.master = new_master.contiguous()
        if self.grad is not None:
            self.grad = self.grad.to(device=self.master.device)
        self.shadow = self.shadow.to(device=self.master.device)
        self.shadow_version = -1

    def set_compute_dtype(self, dtype: torch.dtype) -> None:
        """bf16 (default) or fp16: re-allocates the 16-bit compute copy; the next forward re-casts it from the master."""
        if dtype not in (torch.bfloat16, torch.float16):
            raise TypeError(f"compute dtype must be torch.bfloat16 or torch.float16, not {dtype}")
        if dtype != self.compute_dtype:
            self.compute_dtype = dtype
            self.shadow = torch.zeros(self.numel, dtype=dtype, device=self.master.device)
            self.shadow_version = -1

    def consolidate(self) -> None:
        if self.master_stale and self.master_consolidator is not None:
            self.master_consolidator()

    # ---- shadow maintenance
    def current_version(self) -> int:
        # every in-place edit through torch (load_state_dict, torch optimizers, init) bumps the edited Parameter's
        # version counter; the fused Adam writes through raw pointers and does not.
        return sum(p._version for p in self.params)

    def sync_shadow(self, force: bool = False) -> None:
        """Refresh the bf16 compute copy if the fp32 master was modified through torch (load_state_dict, a torch
        optimizer, manual edits). The fused Adam keeps both in step without bumping any version counter."""
        if self.master is None:  # sharded master: the 16-bit copy is written by the optimizer / the loader, nothing to re-cast
            return
        v = self.current_version()
        if force or v != self.shadow_version:
            # under ZeRO-1 a torch-side edit must rewrite the whole master on every rank (load_state_dict, broadcast);
            # for partial edits call consolidate() first
            self.master_stale = False
            if self.master.is_cuda:
                from . import kernels as K

                K.cast_f32_to_bf16(self.master, self.shadow)
            else:
                self.shadow.copy_(self.master)
            self.shadow_version = v

    def zero_grad(self) -> None:
        if self.grad is not None:
            self.grad.zero_()
        elif self.grad_zero_fn is not None:
            self.grad_zero_fn()
        self.pending_grad_scale = None
