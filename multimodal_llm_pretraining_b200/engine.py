"""Data-parallel step engine: what HF Trainer + accelerate (DDP) / DeepSpeed (ZeRO-1) do for the benchmarked configs,
on top of the flat parameter store (src/train.py:126-171 selects the strategy declaratively in the reference;
src/benchmarking/utils.py:61-80 drives it).

One process per GPU; `torch.distributed` (NCCL over NVLink 5 / NVSwitch) is the plumbing. The path shards naturally:
every rank runs independent micro-batches; the only exchange is per optimizer step.

  strategy "none"  : single GPU.
  strategy "ddp"   : per-layer buckets of the flat fp32 grad buffer are all-reduced (AVG) on a side stream as soon as
                     the layer's backward has produced them (last micro-batch of the accumulation window only), then
                     every rank runs the replicated fused Adam.
  strategy "zero1" : the same buckets are reduce-scattered (AVG) in place — rank r keeps slice r of every bucket —
                     local sum-of-squares + one scalar all-reduce give the global grad norm, the fused Adam updates
                     only the owned slices (moments exist only for them: 8 B/param/W) and writes their bf16 compute
                     copy in the same pass; the bf16 slices (2 B/param — what the GEMMs of the next forward read) are
                     all-gathered in place, and the 1-D parameters (biases, LayerNorm affine: read in fp32 by the kernels)
                     are exchanged in fp32 through one small packed all-reduce. The fp32 master of the 2-D parameters a
                     rank does not own goes stale, exactly as under DeepSpeed
                     ZeRO-1 where the fp32 master exists only on the owner; `consolidate_master()` (called by
                     `state_dict()`) all-gathers it on demand.

Bucket collectives overlapped with backward run on a side stream through a dedicated NCCL communicator capped at
`comm_max_ctas` CTAs: the persistent GEMM grids are sized to the SM count, so every SM a collective occupies delays a whole
wave of tiles; a narrow communicator still moves a layer's gradients well inside that layer's backward time.

`CommPlan` holds the pure bucket/ownership arithmetic and the collective calls so that it can be exercised with gloo on
CPU tensors (tests/test_engine_cpu.py, world_size 2).
"""

from __future__ import annotations

from typing import Callable

import math
import os

import torch
import torch.distributed as dist


class CommPlan:
    """Buckets are [start, end) element ranges of the flat buffers (multiples of 64 elements)."""

    def __init__(self, buckets: list[tuple[int, int]], world_size: int, rank: int, group=None):
        self.buckets = sorted(buckets)
        self.W, self.rank, self.group = world_size, rank, group
        for s, e in self.buckets:
            if (e - s) % (world_size * 4) != 0:
                raise ValueError(f"bucket [{s},{e}) is not divisible into {world_size} 16-byte-aligned slices")

    def owned_slice(self, bucket: tuple[int, int], rank: int | None = None) -> tuple[int, int]:
        s, e = bucket
        r = self.rank if rank is None else rank
        n = (e - s) // self.W
        return s + r * n, s + (r + 1) * n

    def owned_ranges(self, rank: int | None = None) -> list[tuple[int, int]]:
        return [self.owned_slice(b, rank) for b in self.buckets]

    def bucket_of(self, start: int, end: int) -> tuple[int, int]:
        for b in self.buckets:
            if b[0] == start and b[1] == end:
                return b
        raise KeyError((start, end))

    # ---- collectives on one bucket of a flat tensor
    def _gloo(self) -> bool:
        return dist.get_backend(self.group) == "gloo"

    def all_reduce_avg(self, flat: torch.Tensor, bucket: tuple[int, int]) -> None:
        t = flat[bucket[0]:bucket[1]]
        if self._gloo():  # gloo has no AVG
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            t.div_(self.W)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)

    def reduce_scatter_avg(self, flat: torch.Tensor, bucket: tuple[int, int]) -> None:
        """In place: afterwards flat[owned_slice(bucket)] holds the mean over ranks; other slices are undefined."""
        t = flat[bucket[0]:bucket[1]]
        lo, hi = self.owned_slice(bucket)
        if self._gloo():  # test backend: emulate with all-reduce
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            t.div_(self.W)
        else:
            dist.reduce_scatter_tensor(flat[lo:hi], t, op=dist.ReduceOp.AVG, group=self.group)

    def all_gather(self, flat: torch.Tensor, bucket: tuple[int, int]) -> None:
        """In place: every rank contributes its owned slice of the bucket."""
        t = flat[bucket[0]:bucket[1]]
        lo, hi = self.owned_slice(bucket)
        if self._gloo():
            parts = [torch.empty(hi - lo, dtype=flat.dtype, device=flat.device) for _ in range(self.W)]
            dist.all_gather(parts, flat[lo:hi].clone(), group=self.group)
            for r, p_ in enumerate(parts):
                a, b = self.owned_slice(bucket, r)
                flat[a:b].copy_(p_)
        else:
            dist.all_gather_into_tensor(t, flat[lo:hi], group=self.group)

    def all_reduce_sum_scalar(self, x: torch.Tensor) -> None:
        dist.all_reduce(x, op=dist.ReduceOp.SUM, group=self.group)

    def owned_mask(self, idx: torch.Tensor) -> torch.Tensor:
        """bool mask over the flat-buffer element indices `idx`: True where this rank owns the element."""
        own = torch.zeros(int(idx.max().item()) + 1 if idx.numel() else 0, dtype=torch.bool)
        for lo, hi in self.owned_ranges():
            own[lo:hi] = True
        return own[idx.cpu()].to(idx.device)

    def exchange_owned(self, flat: torch.Tensor, idx: torch.Tensor, own: torch.Tensor) -> None:
        """flat[idx] <- the owning rank's value, on every rank: owners contribute their elements to one packed buffer, the
        others zeros, and a SUM all-reduce fills it in (x + 0 is exact)."""
        if idx.numel() == 0:
            return
        vals = flat.index_select(0, idx)
        packed = torch.where(own, vals, torch.zeros_like(vals))
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=self.group)
        flat.index_copy_(0, idx, packed)


class TrainEngine:
    """manual_training_step / manual_optimization_step of the reference harness (src/benchmarking/utils.py:61-80) for a
    B200 module + B200Adam, with DDP or ZeRO-1 over the flat buffers."""

    def __init__(self, model, optimizer, scheduler=None, max_grad_norm: float = 1.0, gradient_accumulation_steps: int = 1,
                 strategy: str = "none", group=None, overlap: bool = True, comm_max_ctas: int | None = None,
                 profile_phases: bool = False):
        self.model, self.optimizer, self.scheduler = model, optimizer, scheduler
        self.max_grad_norm = max_grad_norm
        self.ga = gradient_accumulation_steps
        self.strategy = strategy
        self.flat = model.flat
        self.micro = 0
        self.plan: CommPlan | None = None
        self.comm_stream = None
        self.last_grad_norm = None
        self.overlap = overlap
        self.overlap_group = group
        self.profile_phases = profile_phases
        self.last_phase_ms: dict | None = None
        self._pending: list[tuple[int, int]] = []
        if strategy not in ("none", "ddp", "zero1"):
            raise ValueError(strategy)
        if strategy != "none":
            if not dist.is_initialized():
                raise RuntimeError("strategy %r needs an initialised torch.distributed process group" % strategy)
            W, r = dist.get_world_size(group), dist.get_rank(group)
            self.plan = CommPlan(model.comm_buckets(), W, r, group)
            # host-side tests drive the exchange logic with CPU tensors over gloo: no side stream there
            self.comm_stream = torch.cuda.Stream() if self.flat.master.is_cuda else None
            if comm_max_ctas is None:
                comm_max_ctas = int(os.environ.get("B200_COMM_MAX_CTAS", "0")) or None
            if comm_max_ctas and overlap and dist.get_backend(group) == "nccl":
                opts = dist.ProcessGroupNCCL.Options()
                opts.config.max_ctas = comm_max_ctas
                opts.config.min_ctas = min(comm_max_ctas, 4)
                ranks = dist.get_process_group_ranks(group) if group is not None else list(range(W))
                self.overlap_group = dist.new_group(ranks=ranks, backend="nccl", pg_options=opts)
            self.overlap_plan = CommPlan(model.comm_buckets(), W, r, self.overlap_group)
            if strategy == "zero1":
                optimizer.set_shard(self.plan.owned_ranges())
                self.flat.master_consolidator = self.consolidate_master
                self._build_fp32_exchange()
            # identical initial parameters everywhere (rank 0 wins), like DDP's constructor broadcast
            dist.broadcast(self.flat.master, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            self.flat.sync_shadow(force=True)

    def _build_fp32_exchange(self) -> None:
        """ZeRO-1 replicates the bf16 compute copy, but the kernels read biases and LayerNorm affine parameters (every 1-D
        parameter) from the fp32 master. Those few elements (13 h per layer) are exchanged in fp32 after each optimizer step:
        every rank contributes the elements it owns to one packed buffer, zeros elsewhere, and a SUM all-reduce fills it in."""
        f = self.flat
        dev = f.master.device
        idx = []
        for name in f.names:
            if len(f.shapes[name]) < 2:
                o = f.offsets[name]
                idx.append(torch.arange(o, o + math.prod(f.alloc_shapes[name]), dtype=torch.int64))
        self._fp32_idx = torch.cat(idx).to(dev) if idx else torch.empty(0, dtype=torch.int64, device=dev)
        self._fp32_own = self.plan.owned_mask(self._fp32_idx)

    def _exchange_fp32_params(self) -> None:
        f = self.flat
        self.plan.exchange_owned(f.master, self._fp32_idx, self._fp32_own)
        f.shadow_version = f.current_version()  # a torch-side write to the master that must NOT trigger a shadow re-cast

    # ------------------------------------------------------------------ fwd + bwd of one micro-batch
    def _reduce_bucket(self, plan: CommPlan, b: tuple[int, int]) -> None:
        if self.strategy == "ddp":
            plan.all_reduce_avg(self.flat.grad, b)
        else:
            plan.reduce_scatter_avg(self.flat.grad, b)

    def _on_grads_ready(self, start: int, end: int) -> None:
        b = self.plan.bucket_of(start, end)
        if not self.overlap or self.comm_stream is None:
            self._pending.append(b)
            return
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ev)
            self._reduce_bucket(self.overlap_plan, b)

    def manual_training_step(self, inputs: dict) -> torch.Tensor:
        """One micro-batch forward + backward with gradients ACCUMULATED; the loss is divided by the accumulation count
        before backward (HF:trainer.py:1925-1927). Returns the (undivided) loss as a device scalar."""
        boundary = (self.micro + 1) % self.ga == 0
        self.model.grad_ready_hook = self._on_grads_ready if (self.plan is not None and boundary) else None
        out = self.model(**inputs)
        loss = out["loss"]
        (loss / self.ga if self.ga > 1 else loss).backward()
        self.micro += 1
        return loss.detach()

    # ------------------------------------------------------------------ optimizer step
    def manual_optimization_step(self) -> None:
        """clip (max_grad_norm > 0) -> optimizer.step -> lr_scheduler.step -> zero_grad (src/benchmarking/utils.py:65-80)."""
        from . import kernels as K

        f = self.flat
        marks = [self._mark()]
        if self.plan is not None:
            if self.comm_stream is not None:
                torch.cuda.current_stream().wait_stream(self.comm_stream)
            for b in self._pending:  # overlap=False: the bucket collectives run here, on the compute stream
                self._reduce_bucket(self.plan, b)
            self._pending.clear()
        marks.append(self._mark())
        if self.max_grad_norm is not None and self.max_grad_norm > 0:
            sumsq = torch.zeros((), dtype=torch.float32, device=f.grad.device)
            if self.strategy == "zero1":
                for lo, hi in self.plan.owned_ranges():
                    K.sumsq_(f.grad[lo:hi], sumsq)
                self.plan.all_reduce_sum_scalar(sumsq)
            else:
                K.sumsq_(f.grad, sumsq)
            norm, coef = K.clip_coef(sumsq, self.max_grad_norm)
            f.pending_grad_scale = coef
            self.last_grad_norm = norm
        marks.append(self._mark())
        self.optimizer.step()
        marks.append(self._mark())
        if self.strategy == "zero1":
            for b in self.plan.buckets:
                self.plan.all_gather(f.shadow, b)
            self._exchange_fp32_params()
            f.master_stale = True  # of the 2-D parameters a rank does not own; nothing on the step path reads those
        marks.append(self._mark())
        if self.scheduler is not None:
            self.scheduler.step()
        self.model.zero_grad()
        marks.append(self._mark())
        if self.profile_phases:
            torch.cuda.synchronize()
            names = ["wait_comm", "grad_norm", "adam", "all_gather", "zero_grad"]
            self.last_phase_ms = {n: marks[i].elapsed_time(marks[i + 1]) for i, n in enumerate(names)}

    def _mark(self):
        if not self.profile_phases:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    # ------------------------------------------------------------------ checkpoint / resume
    def save_checkpoint(self, directory) -> None:
        """HF-Trainer-style checkpoint directory: `pytorch_model.bin` (state_dict with HF key names, written by rank 0 after
        the fp32 master is consolidated), `optimizer.pt` (replicated) or `optimizer_rank{r}.pt` (ZeRO-1: every rank saves the
        moments of the slices it owns), `scheduler.pt`, `trainer_state.json`. Collective: all ranks call it."""
        import json
        from pathlib import Path

        d = Path(directory)
        rank = self.plan.rank if self.plan is not None else 0
        if rank == 0:
            d.mkdir(parents=True, exist_ok=True)
        if self.plan is not None:
            dist.barrier(group=self.plan.group)
        sd = self.model.state_dict()  # consolidates the master under ZeRO-1 (collective)
        if rank == 0:
            torch.save({k: v.detach().cpu() for k, v in sd.items()}, d / "pytorch_model.bin")
            if self.scheduler is not None:
                torch.save(self.scheduler.state_dict(), d / "scheduler.pt")
            (d / "trainer_state.json").write_text(json.dumps({
                "global_step": self.micro // self.ga, "micro_step": self.micro, "strategy": self.strategy,
                "world_size": self.plan.W if self.plan is not None else 1, "gradient_accumulation_steps": self.ga}))
        if self.strategy == "zero1":
            torch.save(self.optimizer.state_dict(), d / f"optimizer_rank{rank}.pt")
        elif rank == 0:
            torch.save(self.optimizer.state_dict(), d / "optimizer.pt")
        if self.plan is not None:
            dist.barrier(group=self.plan.group)

    def load_checkpoint(self, directory) -> None:
        """Inverse of save_checkpoint for the same strategy and world size (optimizer shards are per rank)."""
        import json
        from pathlib import Path

        d = Path(directory)
        state = json.loads((d / "trainer_state.json").read_text())
        W = self.plan.W if self.plan is not None else 1
        if state["strategy"] != self.strategy or state["world_size"] != W:
            raise ValueError(f"checkpoint was written with strategy {state['strategy']!r} on {state['world_size']} rank(s); "
                             f"this engine runs {self.strategy!r} on {W}")
        rank = self.plan.rank if self.plan is not None else 0
        self.model.load_state_dict(torch.load(d / "pytorch_model.bin", map_location="cpu"))
        self.flat.sync_shadow(force=True)
        opt_file = d / (f"optimizer_rank{rank}.pt" if self.strategy == "zero1" else "optimizer.pt")
        self.optimizer.load_state_dict(torch.load(opt_file, map_location="cpu"))
        if self.scheduler is not None and (d / "scheduler.pt").exists():
            self.scheduler.load_state_dict(torch.load(d / "scheduler.pt", map_location="cpu"))
        self.micro = int(state["micro_step"])
        self.model.zero_grad()

    def consolidate_master(self) -> None:
        """ZeRO-1: bring the fp32 master of every slice up to date on every rank (collective; all ranks must call it).
        The owner's fp32 values are authoritative; between optimizer steps only the bf16 compute copy is replicated."""
        f = self.flat
        if self.strategy == "zero1" and getattr(f, "master_stale", False):
            for b in self.plan.buckets:
                self.plan.all_gather(f.master, b)
            f.master_stale = False
            f.shadow_version = f.current_version()  # the bf16 copy is already bf16(master) everywhere
