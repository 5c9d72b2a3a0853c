"""Data-parallel step engine: what HF Trainer + accelerate (DDP) / DeepSpeed (ZeRO-1, ZeRO-2, ZeRO-3) do for the benchmarked configs,
on top of the flat parameter store (src/train.py:126-181 selects the strategy declaratively in the reference;
src/benchmarking/utils.py:61-80 drives it).

One process per GPU; `torch.distributed` (NCCL over NVLink 5 / NVSwitch) is the plumbing. The path shards naturally:
every rank runs independent micro-batches; the only exchange is per optimizer step (per micro-batch under ZeRO-2).

  strategy "none"  : single GPU.
  strategy "ddp"   : per-layer buckets of the flat fp32 grad buffer are all-reduced (AVG) on a side stream as soon as
                     the layer's backward has produced them (last micro-batch of the accumulation window only), then
                     every rank runs the replicated fused Adam. The grad norm is a deterministic reduction, so replicas
                     holding bit-identical gradients take bit-identical steps.
  strategy "zero1" : the same buckets are reduce-scattered (AVG) in place — rank r keeps slice r of every bucket —
                     sum of squares over the owned slices + one scalar all-reduce give the global grad norm, the fused Adam
                     updates only the owned slices (moments exist only for them: 8 B/param/W) and writes their 16-bit compute
                     copy in the same pass; the 16-bit slices (2 B/param — what the GEMMs of the next forward read) are
                     all-gathered in place ON THE SIDE STREAM, bucket by bucket in forward order, and the next forward waits
                     per bucket right before the layer that reads it (`param_wait_hook`), so the gather overlaps zero_grad,
                     the host work between steps and the first layers. The 1-D parameters (biases, LayerNorm affine: read in
                     fp32 by the kernels) are exchanged in fp32 through one small packed all-reduce ahead of the gathers. The
                     fp32 master of the 2-D parameters a rank does not own goes stale, exactly as under DeepSpeed ZeRO-1
                     where the fp32 master exists only on the owner; `consolidate_master()` (called by `state_dict()`)
                     all-gathers it on demand.
  strategy "zero2" : gradient sharding (DeepSpeed stage 2 / FSDP shard_grad_op, src/train.py:126-136,172-181). There is NO
                     full gradient buffer: a layer's backward writes into a transient bucket buffer (a ring of two per bucket
                     size), EVERY micro-batch the bucket is reduce-scattered (AVG) on the side stream as soon as the layer is
                     done, the owned slice is accumulated into a packed fp32 shard accumulator (4 B/param/W) and the buffer
                     is cleared for reuse. The optimizer step then runs on the shard accumulator (packed like the moments).

  strategy "zero3" : parameter sharding on top (DeepSpeed stage 3 / FSDP full_shard, src/train.py:126-136,183-194): the fp32 master
                     AND the 16-bit compute copy exist only as the slices a rank owns (18 B/param/W with moments and the gradient
                     shard); the module announces each bucket right before it reads it, in forward and again in backward, and the
                     engine all-gathers it into a transient buffer (ring of two per bucket size) while prefetching the next one of
                     that direction on the side stream. Exercised with the real module over gloo on CPU
                     (tests/test_host_schedule_cpu.py: equal to ZeRO-2 bit for bit, checkpoint round trip, activation
                     checkpointing); NOT yet run on hardware, hence not selectable through the reference-facing `sharding=zero_3`.

fp16 (the reference's precision for every Pythia but 1b and for RoBERTa): `LossScaler` keeps a dynamic loss scale on the
device; the cross entropy folds it into dlogits, the clip-coefficient kernel divides it out again and reports overflow
(non-finite grad norm); an overflow step is skipped and the scale backs off (torch.amp.GradScaler / DeepSpeed
DynamicLossScaler semantics, src/train.py:143-150).

Bucket collectives overlapped with backward run on a side stream (optionally through a dedicated NCCL communicator capped at
`comm_max_ctas` CTAs). `CommPlan` holds the pure bucket/ownership arithmetic and the collective calls so that it can be
exercised with gloo on CPU tensors (tests/test_engine_cpu.py, world_size 2).
"""

from __future__ import annotations

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

    def reduce_scatter_avg_buffer(self, buf: torch.Tensor) -> torch.Tensor:
        """Same on a stand-alone bucket buffer (ZeRO-2 transient buffers); returns the view of the slice this rank owns."""
        self.reduce_scatter_avg(buf, (0, buf.numel()))
        n = buf.numel() // self.W
        return buf[self.rank * n:(self.rank + 1) * n]

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

    def all_gather_buffer(self, buf: torch.Tensor, own: torch.Tensor) -> None:
        """buf (a whole bucket, W * own.numel() elements) <- the slices every rank owns, in rank order (ZeRO-3 weight gather)."""
        if self._gloo():
            parts = [torch.empty_like(own) for _ in range(self.W)]
            dist.all_gather(parts, own.clone(), group=self.group)
            n = own.numel()
            for r, p_ in enumerate(parts):
                buf[r * n:(r + 1) * n].copy_(p_)
        else:
            dist.all_gather_into_tensor(buf, own, group=self.group)

    def all_reduce_sum_scalar(self, x: torch.Tensor) -> None:
        dist.all_reduce(x, op=dist.ReduceOp.SUM, group=self.group)

    def owned_mask(self, idx: torch.Tensor) -> torch.Tensor:
        """bool mask over the flat-buffer element indices `idx`: True where this rank owns the element."""
        own = torch.zeros(idx.numel(), dtype=torch.bool, device=idx.device)
        for lo, hi in self.owned_ranges():  # one comparison pair per bucket over the (few) indices: no flat-buffer-sized temporary
            own |= (idx >= lo) & (idx < hi)
        return own

    def exchange_owned(self, flat: torch.Tensor, idx: torch.Tensor, own: torch.Tensor) -> None:
        """flat[idx] <- the owning rank's value, on every rank: owners contribute their elements to one packed buffer, the
        others zeros, and a SUM all-reduce fills it in (x + 0 is exact)."""
        if idx.numel() == 0:
            return
        vals = flat.index_select(0, idx)
        packed = torch.where(own, vals, torch.zeros_like(vals))
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=self.group)
        flat.index_copy_(0, idx, packed)


class LossScaler:
    """Dynamic loss scale for fp16 runs, state on the device (kernels.loss_scale_update).

    kind="torch"     : torch.amp.GradScaler defaults (what HF Trainer uses for fp16 without DeepSpeed): 2^16, x2 every 2000
                       clean steps, /2 on every overflow.
    kind="deepspeed" : the reference's DeepSpeed fp16 block (src/train.py:143-150): initial_scale_power 16, loss_scale_window
                       1000, hysteresis 2, min_loss_scale 1."""

    def __init__(self, device, kind: str = "torch", init_scale: float | None = None, growth_factor: float = 2.0,
                 backoff_factor: float = 0.5, growth_interval: int | None = None, min_scale: float | None = None,
                 hysteresis: int | None = None):
        ds = kind == "deepspeed"
        self.growth_factor, self.backoff_factor = growth_factor, backoff_factor
        self.growth_interval = growth_interval if growth_interval is not None else (1000 if ds else 2000)
        self.min_scale = min_scale if min_scale is not None else (1.0 if ds else 0.0)
        self.hysteresis = hysteresis if hysteresis is not None else (2 if ds else 1)
        self.scale = torch.full((1,), float(init_scale if init_scale is not None else 2.0 ** 16), dtype=torch.float32, device=device)
        self.growth_tracker = torch.zeros(1, dtype=torch.int32, device=device)
        self.hysteresis_left = torch.full((1,), self.hysteresis, dtype=torch.int32, device=device)
        self.found_inf = torch.zeros(1, dtype=torch.int32, device=device)
        self.skipped_steps = 0

    def update(self) -> None:
        from . import kernels as K

        K.loss_scale_update(self.scale, self.growth_tracker, self.hysteresis_left, self.found_inf, self.growth_factor,
                            self.backoff_factor, self.growth_interval, self.min_scale, self.hysteresis)

    def state_dict(self) -> dict:
        return {"scale": float(self.scale.item()), "growth_tracker": int(self.growth_tracker.item()),
                "hysteresis_left": int(self.hysteresis_left.item()), "skipped_steps": self.skipped_steps}

    def load_state_dict(self, sd: dict) -> None:
        self.scale.fill_(float(sd["scale"]))
        self.growth_tracker.fill_(int(sd["growth_tracker"]))
        self.hysteresis_left.fill_(int(sd["hysteresis_left"]))
        self.skipped_steps = int(sd.get("skipped_steps", 0))


class TrainEngine:
    """manual_training_step / manual_optimization_step of the reference harness (src/benchmarking/utils.py:61-80) for a
    B200 module + B200Adam, with DDP, ZeRO-1 or ZeRO-2 over the flat buffers."""

    STRATEGIES = ("none", "ddp", "zero1", "zero2", "zero3")
    ZERO = ("zero1", "zero2", "zero3")          # optimizer state sharded
    GRAD_SHARDED = ("zero2", "zero3")           # + gradients sharded (reduce-scatter every micro-batch)

    def __init__(self, model, optimizer, scheduler=None, max_grad_norm: float = 1.0, gradient_accumulation_steps: int = 1,
                 strategy: str = "none", group=None, overlap: bool = True, comm_max_ctas: int | None = None,
                 profile_phases: bool = False, loss_scaler: LossScaler | None = None, overlap_param_gather: bool = True,
                 shard_master: bool = False):
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
        self.overlap_param_gather = overlap_param_gather
        self.overlap_group = group
        self.profile_phases = profile_phases
        self.last_phase_ms: dict | None = None
        self._pending: list[tuple[int, int]] = []
        self._param_events: dict[tuple[int, int], object] = {}
        self._sumsq_partials = None
        if strategy not in self.STRATEGIES:
            raise ValueError(strategy)
        f = self.flat
        # fp16: a loss scale is mandatory (dlogits / n_valid underflow in half precision without it)
        if getattr(f, "compute_dtype", torch.bfloat16) == torch.float16 and loss_scaler is None and f.device.type == "cuda":
            loss_scaler = LossScaler(f.device, kind="deepspeed" if strategy in self.ZERO else "torch")
        self.loss_scaler = loss_scaler
        if loss_scaler is not None:
            model.loss_scale = loss_scaler.scale
        if strategy == "zero3" and not getattr(model, "supports_weight_sharding", False):
            raise NotImplementedError(f"strategy 'zero3' needs a module whose backward announces its weight reads (param_bwd_hook); "
                                      f"{type(model).__name__} does not")
        if strategy != "none":
            if not dist.is_initialized():
                raise RuntimeError("strategy %r needs an initialised torch.distributed process group" % strategy)
            W, r = dist.get_world_size(group), dist.get_rank(group)
            self.plan = CommPlan(model.comm_buckets(), W, r, group)
            # host-side tests drive the exchange logic with CPU tensors over gloo: no side stream there
            self.comm_stream = torch.cuda.Stream() if f.device.type == "cuda" else None
            if comm_max_ctas is None:
                comm_max_ctas = int(os.environ.get("B200_COMM_MAX_CTAS", "0")) or None
            if comm_max_ctas and overlap and dist.get_backend(group) == "nccl":
                opts = dist.ProcessGroupNCCL.Options()
                opts.config.max_ctas = comm_max_ctas
                opts.config.min_ctas = min(comm_max_ctas, 4)
                ranks = dist.get_process_group_ranks(group) if group is not None else list(range(W))
                self.overlap_group = dist.new_group(ranks=ranks, backend="nccl", pg_options=opts)
            self.overlap_plan = CommPlan(model.comm_buckets(), W, r, self.overlap_group)
            if strategy in self.ZERO:
                optimizer.set_shard(self.plan.owned_ranges())
                f.master_consolidator = self.consolidate_master
                self._build_fp32_exchange()
                model.param_wait_hook = self._wait_params
            # identical initial parameters everywhere (rank 0 wins), like DDP's constructor broadcast
            dist.broadcast(f.master, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            f.sync_shadow(force=True)
            if strategy in self.GRAD_SHARDED:
                self._setup_zero2()
            # ZeRO-3 shards the 16-bit weights; the nn.Parameters then cannot stay views of a full fp32 master either
            self.shard_master = (bool(shard_master) or strategy == "zero3") and strategy in self.ZERO
            if self.shard_master:
                self._shard_master()
            if strategy == "zero3":
                self._setup_zero3()
        else:
            self.shard_master = False

    # ------------------------------------------------------------------ ZeRO-1/2: replicated fp32 1-D parameters
    def _build_fp32_exchange(self) -> None:
        """ZeRO replicates the 16-bit compute copy, but the kernels read biases and LayerNorm affine parameters (every 1-D
        parameter) from the fp32 master. Those few elements (13 h per layer) are exchanged in fp32 after each optimizer step:
        every rank contributes the elements it owns to one packed buffer, zeros elsewhere, and a SUM all-reduce fills it in."""
        f = self.flat
        dev = f.device
        idx = []
        for name in f.names:
            if len(f.shapes[name]) < 2:
                o = f.offsets[name]
                idx.append(torch.arange(o, o + math.prod(f.alloc_shapes[name]), dtype=torch.int64))
        self._fp32_idx = torch.cat(idx).to(dev) if idx else torch.empty(0, dtype=torch.int64, device=dev)
        self._fp32_own = self.plan.owned_mask(self._fp32_idx)

    def _exchange_fp32_params(self) -> None:
        f = self.flat
        if f.master is None:  # sharded master: owners contribute from their packed fp32 shard, everyone receives into `small`
            packed = torch.zeros_like(f.small)
            packed[self._own_small_pos] = self.optimizer._p32[self._own_packed_idx]
            dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=self.plan.group)
            f.small.copy_(packed)
            return
        self.plan.exchange_owned(f.master, self._fp32_idx, self._fp32_own)
        f.shadow_version = f.current_version()  # a torch-side write to the master that must NOT trigger a shadow re-cast

    # ------------------------------------------------------------------ ZeRO with a sharded fp32 master (opt-in)
    def _shard_master(self) -> None:
        """True ZeRO partition of the fp32 weights (12 B/param/W of optimizer state per rank instead of 4 + 8/W): the optimizer adopts
        the owned slices as a packed fp32 buffer, the full master is freed, the nn.Parameters become views of the 16-bit copy and the
        1-D parameters the kernels read in fp32 move to the compact replicated `flat.small`."""
        f = self.flat
        p32 = self.optimizer.adopt_master_shard()
        dev = f.device
        # flat offset -> packed offset of this rank's owned elements, for the 1-D parameters only
        own_flat = self._fp32_idx[self._fp32_own]
        packed_off = torch.empty_like(own_flat)
        acc = 0
        for lo, hi in self.plan.owned_ranges():
            m = (own_flat >= lo) & (own_flat < hi)
            packed_off[m] = own_flat[m] - lo + acc
            acc += hi - lo
        f.drop_master(dict(self.model.named_parameters()))
        # position inside `small` of every 1-D element (small keeps the order of the flat store, padded per name like it)
        pos = []
        for n in f.small_names():
            k = math.prod(f.alloc_shapes[n])
            pos.append(torch.arange(f.small_offsets[n], f.small_offsets[n] + k, dtype=torch.int64))
        small_pos = torch.cat(pos).to(dev) if pos else torch.empty(0, dtype=torch.int64, device=dev)
        self._own_small_pos = small_pos[self._fp32_own]
        self._own_packed_idx = packed_off
        f.master_materializer = self._materialize_master
        f.master_loader = self._load_full_master
        assert p32 is self.optimizer._p32

    def _materialize_master(self) -> torch.Tensor:
        """Collective: a fresh full fp32 parameter vector gathered from the owners' shards (state_dict / checkpoints)."""
        f = self.flat
        self.sync_params()
        full = torch.zeros(f.numel, dtype=torch.float32, device=f.device)
        off = 0
        for b, (lo, hi) in zip(self.plan.buckets, self.plan.owned_ranges()):
            full[lo:hi].copy_(self.optimizer._p32[off:off + hi - lo])
            off += hi - lo
            self.plan.all_gather(full, b)
        return full

    def _load_full_master(self, full: torch.Tensor) -> None:
        """Inverse: scatter a full fp32 vector (identical on every rank) into the shard, `small` and the 16-bit copy."""
        f = self.flat
        self.sync_params()
        off = 0
        for lo, hi in self.plan.owned_ranges():
            self.optimizer._p32[off:off + hi - lo].copy_(full[lo:hi])
            off += hi - lo
        for n in f.small_names():
            k = math.prod(f.alloc_shapes[n])
            f.small[f.small_offsets[n]:f.small_offsets[n] + k].copy_(full[f.offsets[n]:f.offsets[n] + k])
        if self.strategy == "zero3":  # only the owned 16-bit slices exist
            for b, (lo, hi) in zip(self.plan.buckets, self.plan.owned_ranges()):
                self._w16[self._w16_off[b]:self._w16_off[b] + hi - lo].copy_(full[lo:hi])
            self._invalidate_weights()
        elif f.device.type == "cuda":
            from . import kernels as K

            K.cast_f32_to_bf16(full, f.shadow)
        else:
            f.shadow.copy_(full)
        self._param_events = {}

    # ------------------------------------------------------------------ ZeRO-2: transient bucket buffers + shard accumulator
    def _setup_zero2(self) -> None:
        f = self.flat
        dev = f.device
        owned = self.plan.owned_ranges()
        self._gshard = torch.zeros(sum(hi - lo for lo, hi in owned), dtype=torch.float32, device=dev)
        self._gshard_off, acc = {}, 0  # packed exactly like the optimizer's moments: owned ranges back to back
        for b, (lo, hi) in zip(self.plan.buckets, owned):
            self._gshard_off[b] = acc
            acc += hi - lo
        by_size: dict[int, int] = {}
        for s, e in self.plan.buckets:
            by_size[e - s] = by_size.get(e - s, 0) + 1
        # two buffers per bucket size that occurs more than once (layers: bucket i+1 is written while bucket i is reduced)
        self._ring = {n: [torch.zeros(n, dtype=torch.float32, device=dev) for _ in range(2 if cnt > 1 else 1)] for n, cnt in by_size.items()}
        self._ring_next = {n: 0 for n in by_size}
        self._ring_free: dict[int, object] = {}   # id(buffer) -> event after which it is zeroed and reusable
        self._active: dict[tuple[int, int], torch.Tensor] = {}
        self._name_bucket = {}
        for name in f.names:
            o = f.offsets[name]
            self._name_bucket[name] = next(b for b in self.plan.buckets if b[0] <= o < b[1])
        f.grad = None  # no full gradient buffer under ZeRO-2
        f.grad_router = self._route_grad
        f.grad_zero_fn = self._gshard.zero_
        for p in f.params:
            p.grad = None

    def zero2_transient_bytes(self) -> int:
        return sum(b.numel() * 4 for ring in self._ring.values() for b in ring)

    def _route_grad(self, name: str):
        """FlatParams.gview*: (buffer, element offset) where the gradient of `name` is accumulated during this backward."""
        b = self._name_bucket[name]
        buf = self._active.get(b)
        if buf is None:
            n = b[1] - b[0]
            ring = self._ring[n]
            buf = ring[self._ring_next[n] % len(ring)]
            self._ring_next[n] += 1
            ev = self._ring_free.pop(id(buf), None)
            if ev is not None:
                torch.cuda.current_stream().wait_event(ev)  # its previous reduce-scatter + clear have finished
            self._active[b] = buf
        return buf, self.flat.offsets[name] - b[0]

    def _reduce_bucket_zero2(self, plan: CommPlan, b: tuple[int, int], buf: torch.Tensor) -> None:
        own = plan.reduce_scatter_avg_buffer(buf)
        off = self._gshard_off[b]
        self._gshard[off:off + own.numel()].add_(own)
        buf.zero_()

    # ------------------------------------------------------------------ ZeRO-3: sharded 16-bit weights, gathered per bucket
    def _setup_zero3(self) -> None:
        """DeepSpeed stage 3 / FSDP full_shard (src/train.py:126-136,183-194): on top of the ZeRO-2 gradient handling and the sharded
        fp32 master, the 16-bit compute copy is sharded too — each rank keeps the slices it owns (2 B/param/W, packed like the moments).
        The module announces every bucket right before it reads it (forward: param_wait_hook, backward: param_bwd_hook); the engine
        all-gathers that bucket into a transient buffer (ring of two per bucket size) and prefetches the next one of the same direction
        on the side stream, so the gather of layer i+1 runs under the kernels of layer i. Per rank: 2 + 4 + 8 + 4 = 18 B/param/W
        (weights, fp32 master, moments, gradient shard) + the transient buffers + the replicated fp32 1-D parameters."""
        f = self.flat
        dev, dtype = f.device, f.shadow.dtype
        owned = self.plan.owned_ranges()
        self._w16 = torch.empty(sum(hi - lo for lo, hi in owned), dtype=dtype, device=dev)
        self._w16_off, acc = {}, 0
        for b, (lo, hi) in zip(self.plan.buckets, owned):
            self._w16_off[b] = acc
            self._w16[acc:acc + hi - lo].copy_(f.shadow[lo:hi])
            acc += hi - lo
        by_size: dict[int, int] = {}
        for s, e in self.plan.buckets:
            by_size[e - s] = by_size.get(e - s, 0) + 1
        self._wring = {n: [torch.empty(n, dtype=dtype, device=dev) for _ in range(2 if cnt > 1 else 1)] for n, cnt in by_size.items()}
        self._wring_next = {n: 0 for n in by_size}
        self._wres: dict[tuple[int, int], torch.Tensor] = {}      # bucket -> buffer the compute stream may read now
        self._wflight: dict[tuple[int, int], tuple] = {}          # bucket -> (buffer, event of its gather | None)
        bwd = [tuple(b) for b in self.model.comm_buckets()]       # head, layers L-1 .. 0, input embedding
        self._w_fwd_order = list(reversed(bwd))
        self._w_bwd_order = bwd[:-1]                              # the embedding's backward is a scatter-add: no weight read
        deps = getattr(self.model, "weight_bucket_deps", None)    # buckets read together (RoBERTa: head + the tied decoder matrix)
        self._w_deps = {tuple(k): [tuple(x) for x in v] for k, v in (deps() if deps else {}).items()}
        f.drop_shadow(dict(self.model.named_parameters()))
        f.weight_router = self._route_weight
        self.model.param_wait_hook = self._need_weights_fwd
        self.model.param_bwd_hook = self._need_weights_bwd

    def zero3_transient_bytes(self) -> int:
        return sum(b.numel() * b.element_size() for ring in self._wring.values() for b in ring)

    def _route_weight(self, name: str):
        """FlatParams.wview*: (buffer, element offset) of the gathered 16-bit copy of `name`."""
        b = self._name_bucket[name]
        buf = self._wres.get(b)
        if buf is None:
            raise RuntimeError(f"ZeRO-3: the weights of {name!r} are not resident — the module read them outside its parameter hooks")
        return buf, self.flat.offsets[name] - b[0]

    def _fetch_weights(self, b: tuple[int, int]) -> None:
        """Start the all-gather of bucket b into the next ring buffer of its size."""
        n = b[1] - b[0]
        ring = self._wring[n]
        buf = ring[self._wring_next[n] % len(ring)]
        self._wring_next[n] += 1
        for d in (self._wres, self._wflight):  # whatever still claimed this buffer is gone once it is overwritten
            for ob in [ob for ob, v in d.items() if (v if d is self._wres else v[0]) is buf]:
                del d[ob]
        off = self._w16_off[b]
        own = self._w16[off:off + n // self.plan.W]
        if self.comm_stream is None or not self.overlap:
            self.plan.all_gather_buffer(buf, own)
            self._wflight[b] = (buf, None)
            return
        # the side stream waits for the compute stream as of NOW: every kernel that read this buffer's previous contents was enqueued
        # before its bucket was released (which happens before any fetch), and so was the cast that last wrote the 16-bit shard
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ev)
            self.overlap_plan.all_gather_buffer(buf, own)
            done = torch.cuda.Event()
            done.record()
        self._wflight[b] = (buf, done)

    def _need_weights(self, start: int, end: int, order: list) -> None:
        b = (start, end)
        need = [b] + self._w_deps.get(b, [])
        for ob in [ob for ob in self._wres if ob not in need]:  # every kernel of the previous buckets is already enqueued: release them
            del self._wres[ob]
        for nb in need:
            if nb not in self._wres:
                if nb not in self._wflight:
                    self._fetch_weights(nb)
                buf, ev = self._wflight.pop(nb)
                if ev is not None:
                    torch.cuda.current_stream().wait_event(ev)
                self._wres[nb] = buf
        i = order.index(b) if b in order else -1
        if 0 <= i < len(order) - 1 and order[i + 1] not in self._wflight and order[i + 1] not in self._wres:
            self._fetch_weights(order[i + 1])

    def _need_weights_fwd(self, start: int, end: int) -> None:
        self._need_weights(start, end, self._w_fwd_order)

    def _need_weights_bwd(self, start: int, end: int) -> None:
        self._need_weights(start, end, self._w_bwd_order)

    def _invalidate_weights(self) -> None:
        """The 16-bit shard changed (optimizer step, load): gathered copies are stale."""
        self._wres.clear()
        self._wflight.clear()

    # ------------------------------------------------------------------ fwd + bwd of one micro-batch
    def _reduce_bucket(self, plan: CommPlan, b: tuple[int, int]) -> None:
        if self.strategy == "ddp":
            plan.all_reduce_avg(self.flat.grad, b)
        else:
            plan.reduce_scatter_avg(self.flat.grad, b)

    def _on_grads_ready(self, start: int, end: int) -> None:
        b = self.plan.bucket_of(start, end)
        if self.strategy in self.GRAD_SHARDED:
            buf = self._active.pop(b, None)
            if buf is None:  # no gradient of this bucket was touched in this backward
                return
            if not self.overlap or self.comm_stream is None:
                self._reduce_bucket_zero2(self.plan, b, buf)
                return
            ev = torch.cuda.Event()
            ev.record()
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ev)
                self._reduce_bucket_zero2(self.overlap_plan, b, buf)
                done = torch.cuda.Event()
                done.record()
            self._ring_free[id(buf)] = done
            return
        if not self.overlap or self.comm_stream is None:
            self._pending.append(b)
            return
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ev)
            self._reduce_bucket(self.overlap_plan, b)

    def _wait_params(self, start: int, end: int) -> None:
        """model.param_wait_hook: called by the forward right before it first reads the parameters of bucket [start, end)."""
        ev = self._param_events.get((start, end))
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)

    def manual_training_step(self, inputs: dict) -> torch.Tensor:
        """One micro-batch forward + backward with gradients ACCUMULATED; the loss is divided by the accumulation count
        before backward (HF:trainer.py:1925-1927). Returns the (undivided, unscaled) loss as a device scalar."""
        boundary = (self.micro + 1) % self.ga == 0
        every = self.strategy in self.GRAD_SHARDED  # gradient sharding reduces every micro-batch
        self.model.grad_ready_hook = self._on_grads_ready if (self.plan is not None and (boundary or every)) else None
        out = self.model(**inputs)
        loss = out["loss"]
        (loss / self.ga if self.ga > 1 else loss).backward()
        self.micro += 1
        return loss.detach()

    # ------------------------------------------------------------------ optimizer step
    def _grad_sumsq(self) -> torch.Tensor:
        from . import kernels as K

        f = self.flat
        sumsq = torch.zeros((), dtype=torch.float32, device=f.device)
        if self.strategy in self.GRAD_SHARDED:
            K.sumsq_(self._gshard, sumsq)
            self.plan.all_reduce_sum_scalar(sumsq)
        elif self.strategy == "zero1":
            if hasattr(self.optimizer, "chunk_table"):  # one launch pair over every owned slice
                cs, cl = self.optimizer.chunk_table()
                if self._sumsq_partials is None or self._sumsq_partials.numel() < cs.numel():
                    self._sumsq_partials = torch.empty(max(cs.numel(), 1), dtype=torch.float32, device=f.grad.device)
                K.sumsq_chunks_(f.grad, cs, cl, sumsq, self._sumsq_partials)
            else:
                for lo, hi in self.plan.owned_ranges():
                    K.sumsq_(f.grad[lo:hi], sumsq)
            self.plan.all_reduce_sum_scalar(sumsq)
        else:
            # deterministic reduction: DDP replicas hold bit-identical all-reduced gradients and must compute the same norm
            K.sumsq_(f.grad, sumsq)
        return sumsq

    def manual_optimization_step(self) -> bool:
        """clip (max_grad_norm > 0) -> optimizer.step -> lr_scheduler.step -> zero_grad (src/benchmarking/utils.py:65-80).
        Returns False when the step was skipped because fp16 gradients overflowed (loss scale backed off), else True."""
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
        scaler = self.loss_scaler
        clip = self.max_grad_norm is not None and self.max_grad_norm > 0
        overflow = False
        if clip or scaler is not None:
            sumsq = self._grad_sumsq()
            if scaler is not None:
                norm, coef = K.clip_coef(sumsq, self.max_grad_norm if clip else 0.0, loss_scale=scaler.scale, found_inf=scaler.found_inf)
                # like torch.amp.GradScaler.step / DeepSpeed: the overflow decision is taken on the host (one 4-byte read per
                # optimizer step); every rank sees the same flag (all-reduced sumsq under ZeRO, identical gradients under DDP)
                overflow = bool(scaler.found_inf.item())
            else:
                norm, coef = K.clip_coef(sumsq, self.max_grad_norm)
            f.pending_grad_scale = coef
            self.last_grad_norm = norm
        marks.append(self._mark())
        if overflow:
            scaler.skipped_steps += 1
            f.pending_grad_scale = None
        elif self.strategy in self.GRAD_SHARDED:
            self.optimizer.step(grads=self._gshard, grads_packed=True)
        else:
            self.optimizer.step()
        if scaler is not None:
            scaler.update()
        marks.append(self._mark())
        if self.strategy in self.ZERO and not overflow:
            self._gather_params()
        marks.append(self._mark())
        if self.scheduler is not None and not overflow:
            self.scheduler.step()
        self.model.zero_grad()
        marks.append(self._mark())
        if self.profile_phases:
            torch.cuda.synchronize()
            names = ["wait_comm", "grad_norm", "adam", "all_gather", "zero_grad"]
            self.last_phase_ms = {n: marks[i].elapsed_time(marks[i + 1]) for i, n in enumerate(names)}
        return not overflow

    def _gather_params(self) -> None:
        """ZeRO: replicate what the optimizer just wrote — fp32 1-D parameters first (small packed all-reduce), then the 16-bit
        compute copy bucket by bucket in FORWARD order. With a side stream the gathers run there and the next forward waits
        per bucket (`_wait_params`), so they overlap zero_grad, the host gap between steps and the first layers."""
        f = self.flat
        if self.strategy == "zero3":  # nothing to replicate: refresh the rank's 16-bit shard from its fp32 shard, and the fp32 1-D parameters
            if f.device.type == "cuda":
                from . import kernels as K

                K.cast_f32_to_bf16(self.optimizer._p32, self._w16)
            else:
                self._w16.copy_(self.optimizer._p32)
            self._exchange_fp32_params()
            self._invalidate_weights()
            return
        fwd_order = list(reversed(self.model.comm_buckets()))  # comm_buckets() is backward order
        side = self.comm_stream is not None and self.overlap_param_gather
        if not side:
            for b in fwd_order:
                self.plan.all_gather(f.shadow, b)
            self._exchange_fp32_params()
            self._param_events = {}
        else:
            ev = torch.cuda.Event()
            ev.record()
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ev)
                self._exchange_fp32_params()
                events = {}
                for b in fwd_order:
                    self.plan.all_gather(f.shadow, b)
                    e = torch.cuda.Event()
                    e.record()
                    events[tuple(b)] = e
            self._param_events = events
        f.master_stale = f.master is not None  # of the 2-D parameters a rank does not own; nothing on the step path reads those

    def sync_params(self) -> None:
        """Make the current stream wait for every outstanding parameter gather (before reading the parameters outside the
        model's own forward: checkpoints, consolidation, evaluation code that bypasses the hooks)."""
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)

    def _mark(self):
        if not self.profile_phases:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    # ------------------------------------------------------------------ checkpoint / resume
    def save_checkpoint(self, directory) -> None:
        """HF-Trainer-style checkpoint directory: `pytorch_model.bin` (state_dict with HF key names, written by rank 0 after
        the fp32 master is consolidated), `optimizer.pt` (replicated) or `optimizer_rank{r}.pt` (ZeRO: every rank saves the
        moments of the slices it owns), `scheduler.pt`, `trainer_state.json` (step counters, the dropout step seed of the
        module, the fp16 loss-scaler state). Collective: all ranks call it."""
        import json
        from pathlib import Path

        d = Path(directory)
        rank = self.plan.rank if self.plan is not None else 0
        if rank == 0:
            d.mkdir(parents=True, exist_ok=True)
        if self.plan is not None:
            dist.barrier(group=self.plan.group)
        self.sync_params()
        sd = self.model.state_dict()  # consolidates the master under ZeRO (collective)
        if rank == 0:
            torch.save({k: v.detach().cpu() for k, v in sd.items()}, d / "pytorch_model.bin")
            if self.scheduler is not None:
                torch.save(self.scheduler.state_dict(), d / "scheduler.pt")
            state = {"global_step": self.micro // self.ga, "micro_step": self.micro, "strategy": self.strategy,
                     "world_size": self.plan.W if self.plan is not None else 1, "gradient_accumulation_steps": self.ga,
                     # RoBERTa's counter-based dropout masks are functions of this counter: a resumed run must not replay them
                     "dropout_step_seed": int(getattr(self.model, "_step_seed", 0)),
                     "loss_scaler": self.loss_scaler.state_dict() if self.loss_scaler is not None else None}
            (d / "trainer_state.json").write_text(json.dumps(state))
        if self.strategy in self.ZERO:
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
        self.sync_params()
        self.model.load_state_dict(torch.load(d / "pytorch_model.bin", map_location="cpu"))
        self.flat.sync_shadow(force=True)
        self._param_events = {}
        opt_file = d / (f"optimizer_rank{rank}.pt" if self.strategy in self.ZERO else "optimizer.pt")
        self.optimizer.load_state_dict(torch.load(opt_file, map_location="cpu"))
        if self.scheduler is not None and (d / "scheduler.pt").exists():
            self.scheduler.load_state_dict(torch.load(d / "scheduler.pt", map_location="cpu"))
        self.micro = int(state["micro_step"])
        if hasattr(self.model, "_step_seed"):
            self.model._step_seed = int(state.get("dropout_step_seed", 0))
        if self.loss_scaler is not None and state.get("loss_scaler"):
            self.loss_scaler.load_state_dict(state["loss_scaler"])
        self.model.zero_grad()

    def consolidate_master(self) -> None:
        """ZeRO: bring the fp32 master of every slice up to date on every rank (collective; all ranks must call it).
        The owner's fp32 values are authoritative; between optimizer steps only the 16-bit compute copy is replicated."""
        f = self.flat
        if f.master is None:
            return  # sharded master: materialised on demand (flat.materialize_master)
        if self.strategy in self.ZERO and getattr(f, "master_stale", False):
            self.sync_params()
            for b in self.plan.buckets:
                self.plan.all_gather(f.master, b)
            f.master_stale = False
            f.shadow_version = f.current_version()  # the 16-bit copy is already cast(master) everywhere
