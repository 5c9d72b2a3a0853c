"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: bucket/ownership arithmetic of CommPlan, the DDP and
ZeRO-1 exchange patterns on flat buffers, and their equivalence to a single-process step with the 2x batch."""
import os
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from multimodal_llm_pretraining_b200.engine import CommPlan  # noqa: E402

BUCKETS = [(0, 256), (256, 1024), (1024, 1152)]
N = 1152


def _adam_ref(p, g, m, v, t, lr=1e-2, b1=0.9, b2=0.95, eps=1e-8):
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    p.addcdiv_(m, v.sqrt() / (1 - b2 ** t) ** 0.5 + eps, value=-lr / (1 - b1 ** t))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = CommPlan(BUCKETS, world, rank)
        torch.manual_seed(0)
        master0 = torch.randn(N)
        grads = [torch.randn(N, generator=torch.Generator().manual_seed(10 + r)) for r in range(world)]
        mean_grad = sum(grads) / world

        # --- DDP: bucketed all-reduce(avg) leaves the mean everywhere
        g = grads[rank].clone()
        for b in plan.buckets:
            plan.all_reduce_avg(g, b)
        assert torch.allclose(g, mean_grad, atol=1e-6)

        # --- ZeRO-1: reduce-scatter -> norm over owned slices + scalar all-reduce -> sharded Adam -> all-gather
        g = grads[rank].clone()
        for b in plan.buckets:
            plan.reduce_scatter_avg(g, b)
        owned = plan.owned_ranges()
        assert sum(hi - lo for lo, hi in owned) == N // world
        for lo, hi in owned:
            assert torch.allclose(g[lo:hi], mean_grad[lo:hi], atol=1e-6)
        sumsq = torch.zeros(())
        for lo, hi in owned:
            sumsq += (g[lo:hi] ** 2).sum()
        plan.all_reduce_sum_scalar(sumsq)
        assert torch.allclose(sumsq, (mean_grad ** 2).sum(), rtol=1e-5)
        coef = min(1.0, 1.0 / (float(sumsq.sqrt()) + 1e-6))
        master = master0.clone()
        m = torch.zeros(N // world)
        v = torch.zeros(N // world)
        off = 0
        for lo, hi in owned:  # moments packed back to back, like B200Adam.set_shard
            _adam_ref(master[lo:hi], g[lo:hi] * coef, m[off:off + hi - lo], v[off:off + hi - lo], 1)
            off += hi - lo
        # what the engine replicates between steps is the bf16 compute copy of the owned slices (2 B/param) ...
        shadow = torch.zeros(N, dtype=torch.bfloat16)
        for lo, hi in owned:
            shadow[lo:hi] = master[lo:hi].to(torch.bfloat16)
        for b in plan.buckets:
            plan.all_gather(shadow, b)
        # ... and the fp32 master only on demand (consolidate_master)
        for b in plan.buckets:
            plan.all_gather(master, b)
        assert torch.equal(shadow, master.to(torch.bfloat16))
        # the fp32 elements the kernels read directly (1-D parameters) travel through one packed SUM all-reduce
        stale = torch.full((N,), float(rank + 7))          # every rank starts with its own (wrong) values ...
        truth = torch.arange(N, dtype=torch.float32) * 0.5
        for lo, hi in owned:
            stale[lo:hi] = truth[lo:hi]                      # ... except the elements it owns
        idx = torch.cat([torch.arange(100, 164), torch.arange(300, 700), torch.arange(1100, 1152)])
        own_mask = plan.owned_mask(idx)
        assert int(own_mask.sum()) == sum(max(0, min(hi, b) - max(lo, a)) for lo, hi in owned for a, b in [(100, 164), (300, 700), (1100, 1152)])
        before = stale.clone()
        plan.exchange_owned(stale, idx, own_mask)
        assert torch.equal(stale[idx], truth[idx])
        rest = torch.ones(N, dtype=torch.bool)
        rest[idx] = False
        assert torch.equal(stale[rest], before[rest])        # nothing else is touched
        # single-process reference with the averaged gradient
        ref = master0.clone()
        _adam_ref(ref, mean_grad * coef, torch.zeros(N), torch.zeros(N), 1)
        assert torch.allclose(master, ref, atol=1e-6), (master - ref).abs().max()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_ddp_and_zero1_exchange_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_commplan_ownership_partitions_every_bucket():
    for W in (1, 2, 4, 8):
        plans = [CommPlan([(0, 512), (512, 4096)], W, r) for r in range(W)]
        cover = torch.zeros(4096, dtype=torch.int32)
        for pl in plans:
            for lo, hi in pl.owned_ranges():
                assert lo % 4 == 0 and hi % 4 == 0
                cover[lo:hi] += 1
        assert torch.all(cover == 1)
    with pytest.raises(ValueError):
        CommPlan([(0, 100)], 8, 0)


def test_model_buckets_cover_flat_buffer_in_backward_order():
    from multimodal_llm_pretraining_b200.modeling_gpt_neox import B200GPTNeoXForCausalLM
    from multimodal_llm_pretraining_b200.models.configs import as_namespace, pythia_config_dict

    cfg = dict(pythia_config_dict("pythia-70m"), num_hidden_layers=3, vocab_size=1024)
    m = B200GPTNeoXForCausalLM(as_namespace(cfg))
    buckets = m.comm_buckets()
    assert len(buckets) == 3 + 2
    srt = sorted(buckets)
    assert srt[0][0] == 0
    for (a, b), (c, d) in zip(srt, srt[1:]):
        assert b == c, "buckets must tile the flat buffer without gaps"
    assert all((b - a) % 64 == 0 for a, b in buckets)
    n_params = sum(p.numel() for p in m.parameters())
    assert srt[-1][1] >= n_params and srt[-1][1] <= m.flat.numel
    # backward order: head first, input embedding last
    assert buckets[0] == m._head_range() and buckets[-1][0] == 0


def test_b200adam_shard_chunk_table_cpu():
    """Chunk table of the ZeRO-1 shard: owned ranges map to packed moment offsets and never cross param boundaries."""
    from multimodal_llm_pretraining_b200.modeling_gpt_neox import B200GPTNeoXForCausalLM
    from multimodal_llm_pretraining_b200.models.configs import as_namespace, pythia_config_dict
    from multimodal_llm_pretraining_b200.optim import B200Adam

    cfg = dict(pythia_config_dict("pythia-70m"), num_hidden_layers=2, vocab_size=512)
    m = B200GPTNeoXForCausalLM(as_namespace(cfg))
    plan = CommPlan(m.comm_buckets(), 4, 1)
    opt = B200Adam(m.parameters(), lr=1e-3)
    opt.set_shard(plan.owned_ranges())
    opt._build()
    starts, lens, soff = opt._chunk_start.tolist(), opt._chunk_len.tolist(), opt._chunk_state.tolist()
    owned = plan.owned_ranges()
    total_owned = sum(hi - lo for lo, hi in owned)
    assert opt._m.numel() == total_owned
    for s, l, so in zip(starts, lens, soff):
        assert any(lo <= s and s + l <= hi for lo, hi in owned)
        assert 0 <= so and so + l <= total_owned
    # chunks cover exactly the parameter elements inside the owned ranges
    covered = sum(lens)
    expect = 0
    for p in m.parameters():
        _, off, n = p._b200_flat
        for lo, hi in owned:
            expect += max(0, min(off + n, hi) - max(off, lo))
    assert covered == expect


# ---------------------------------------------------------------------------------------------------------------
# TrainEngine host logic end to end (bucket hooks, reduce-scatter / all-reduce, sharded optimizer hand-off, bf16 all-gather,
# fp32 exchange of the 1-D parameters, stale-master bookkeeping, checkpoint round trip) over gloo with CPU tensors.
# The CUDA kernels the engine calls (sum of squares, clip coefficient, fused Adam) are replaced by torch statements of the
# same arithmetic IN THE TEST; the toy model reads 2-D parameters from the bf16 compute copy and 1-D parameters from the fp32
# master, exactly like the real modules do (modeling_gpt_neox.py: _w / _p).
# ---------------------------------------------------------------------------------------------------------------
def _toy_classes():
    import torch.nn as nn

    from multimodal_llm_pretraining_b200.flat import FlatParams
    from multimodal_llm_pretraining_b200.modeling_gpt_neox import _FlatModule, _Params

    class _Loss(torch.autograd.Function):
        @staticmethod
        def forward(ctx, model, x, y, anchor):
            f = model.flat
            leaves = {}
            for n in f.names:  # 2-D from the bf16 compute copy, 1-D from the fp32 master
                src = f.view(f.shadow, n).float() if len(f.shapes[n]) == 2 else f.pview(n).clone()
                leaves[n] = src.requires_grad_(True)
            with torch.enable_grad():
                h = torch.tanh(leaves["emb.weight"][x] @ leaves["l0.weight"].t() + leaves["l0.bias"])
                out = h @ leaves["l1.weight"].t() + leaves["l1.bias"]
                loss = ((out - y) ** 2).mean()
            ctx.model, ctx.leaves, ctx.loss = model, leaves, loss
            return loss.detach()

        @staticmethod
        def backward(ctx, grad_out):
            model, f = ctx.model, ctx.model.flat
            names = list(ctx.leaves)
            grads = torch.autograd.grad(ctx.loss, [ctx.leaves[n] for n in names], grad_out)
            scale = getattr(model, "loss_scale", None)  # fp16 runs: the engine's device-side loss scale multiplies every gradient
            for n, g in zip(names, grads):
                f.gview(n).add_(g if scale is None else g * float(scale))  # the flat gradient buffer, or the engine's transient bucket buffer under ZeRO-2
            if model.grad_ready_hook:
                for b in model.comm_buckets():  # backward order: head first, embedding last
                    model.grad_ready_hook(*b)
            return None, None, None, None

    class Toy(_FlatModule):
        def __init__(self, seed=0):
            super().__init__()
            self.flat = FlatParams([("emb.weight", (32, 8)), ("l0.weight", (24, 8)), ("l0.bias", (24,)), ("l1.weight", (4, 24)), ("l1.bias", (4,))])
            self.emb = _Params(self.flat, "emb", ("weight",))
            self.l0 = _Params(self.flat, "l0", ("weight", "bias"))
            self.l1 = _Params(self.flat, "l1", ("weight", "bias"))
            self.grad_ready_hook = None
            g = torch.Generator().manual_seed(seed)
            with torch.no_grad():
                for p in self.parameters():
                    p.copy_(torch.randn(p.shape, generator=g) * 0.3)
            self.flat.sync_shadow(force=True)

        def comm_buckets(self):
            f = self.flat
            return [f.range_of(["l1.weight", "l1.bias"]), f.range_of(["l0.weight", "l0.bias"]), f.range_of(["emb.weight"])]

        def forward(self, input_ids, labels):
            self.flat.sync_shadow()
            return {"loss": _Loss.apply(self, input_ids, labels, self.emb.weight)}

    class ToyAdam:
        """Adam over the owned ranges of the flat store, writing the fp32 master and its bf16 copy (what adam.cu does)."""

        def __init__(self, flat, lr=1e-2):
            self.flat, self.lr, self.t = flat, lr, 0
            self.set_shard([(0, flat.numel)])

        def set_shard(self, ranges):
            self.ranges = [tuple(r) for r in ranges]
            n = sum(hi - lo for lo, hi in self.ranges)
            self._m, self._v = torch.zeros(n), torch.zeros(n)
            self._p32 = None

        def adopt_master_shard(self):  # sharded fp32 master: the owned slices packed like the moments
            self._p32 = torch.cat([self.flat.master[lo:hi] for lo, hi in self.ranges]).clone()
            return self._p32

        def step(self, grads=None, grads_packed=False):
            f = self.flat
            self.t += 1
            scale = 1.0 if f.pending_grad_scale is None else float(f.pending_grad_scale)
            f.pending_grad_scale = None
            off = 0
            for lo, hi in self.ranges:
                n = hi - lo
                g = grads[off:off + n] if grads_packed else f.grad[lo:hi]  # ZeRO-2: packed shard accumulator, laid out like m / v
                pm = f.master[lo:hi] if self._p32 is None else self._p32[off:off + n]
                _adam_ref(pm, g * scale, self._m[off:off + n], self._v[off:off + n], self.t, lr=self.lr)
                f.shadow[lo:hi] = pm.to(torch.bfloat16)
                off += n

        def state_dict(self):
            return {"t": self.t, "m": self._m.clone(), "v": self._v.clone()}

        def load_state_dict(self, sd):
            self.t = sd["t"]
            self._m.copy_(sd["m"]), self._v.copy_(sd["v"])

    return Toy, ToyAdam


def _engine_worker(rank, world, port, q, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import multimodal_llm_pretraining_b200.kernels as K
        from multimodal_llm_pretraining_b200.engine import TrainEngine

        K.sumsq_ = lambda x, out: out.add_((x.double() ** 2).sum().float())
        K.clip_coef = lambda sumsq, max_norm, **kw: (sumsq.sqrt(), torch.clamp(max_norm / (sumsq.sqrt() + 1e-6), max=1.0))
        Toy, ToyAdam = _toy_classes()
        g = torch.Generator().manual_seed(100)
        xs = torch.randint(0, 32, (3, 2, world, 16), generator=g)       # [step, micro, rank, tokens]
        ys = torch.randn(3, 2, world, 16, 4, generator=g)

        def run(eng, steps):
            for s in steps:
                for m in range(2):
                    eng.manual_training_step({"input_ids": xs[s, m, rank], "labels": ys[s, m, rank]})
                eng.manual_optimization_step()

        finals = {}
        for strategy in ("ddp", "zero1"):
            model = Toy()
            eng = TrainEngine(model, ToyAdam(model.flat), None, max_grad_norm=0.05, gradient_accumulation_steps=2, strategy=strategy)
            run(eng, range(3))
            f = model.flat
            if strategy == "zero1":
                assert f.master_stale and eng.optimizer._m.numel() == sum(hi - lo for lo, hi in eng.plan.owned_ranges()) < f.numel
                # between steps: bf16 copy and fp32 1-D parameters are replicated ...
                ref = finals["ddp"]
                assert torch.equal(f.shadow, ref["shadow"])
                assert torch.equal(f.master[eng._fp32_idx], ref["master"][eng._fp32_idx])
                # ... the fp32 master of foreign 2-D slices is not, until consolidation
                assert not torch.equal(f.master, ref["master"])
                sd = model.state_dict()
                assert not f.master_stale and torch.equal(f.master, ref["master"])
                assert set(sd) == {"emb.weight", "l0.weight", "l0.bias", "l1.weight", "l1.bias"}
            finals[strategy] = {"master": f.master.clone(), "shadow": f.shadow.clone()}
        moved = (finals["ddp"]["master"] - Toy().flat.master).abs().max()
        assert moved > 1e-3, "the toy problem must actually train"

        # ZeRO-2: no full gradient buffer; every micro-batch's bucket is reduce-scattered and accumulated into the shard
        model = Toy()
        eng = TrainEngine(model, ToyAdam(model.flat), None, max_grad_norm=0.05, gradient_accumulation_steps=2, strategy="zero2")
        assert model.flat.grad is None and all(p.grad is None for p in model.parameters())
        assert eng._gshard.numel() == sum(hi - lo for lo, hi in eng.plan.owned_ranges()) < model.flat.numel
        assert eng.zero2_transient_bytes() == 4 * sum(e - s for s, e in eng.plan.buckets)  # three distinct bucket sizes here: one buffer each
        run(eng, range(3))
        assert float(eng._gshard.abs().max()) == 0.0 and not eng._active
        model.state_dict()
        # same mean gradient, summed in a different order (per micro-batch across ranks, then over micro-batches)
        ref = finals["ddp"]["master"]
        err = ((model.flat.master - ref).norm() / (ref - Toy().flat.master).norm()).item()
        assert err < 1e-5, err

        # sharded fp32 master (true ZeRO partition of the weights): no full master after construction, parameters are views of the
        # 16-bit copy, fp32 1-D parameters replicated in flat.small; state_dict materialises fp32 from the owners and must equal DDP
        for strategy in ("zero1", "zero2"):
            model = Toy()
            eng = TrainEngine(model, ToyAdam(model.flat), None, max_grad_norm=0.05, gradient_accumulation_steps=2, strategy=strategy, shard_master=True)
            f = model.flat
            assert f.master is None and eng.optimizer._p32.numel() == sum(hi - lo for lo, hi in eng.plan.owned_ranges())
            assert all(p.dtype == torch.bfloat16 and p.grad is None for p in model.parameters())
            run(eng, range(3))
            sd = model.state_dict()
            full = torch.cat([sd[n].reshape(-1) for n in f.names])
            want = torch.cat([f.view(finals["ddp"]["master"], n).reshape(-1) for n in f.names])
            if strategy == "zero1":
                assert torch.equal(full, want), (full - want).abs().max()
            else:
                assert ((full - want).norm() / (want - torch.cat([Toy().flat.view(Toy().flat.master, n).reshape(-1) for n in f.names])).norm()).item() < 1e-5
            assert torch.equal(f.shadow, finals["ddp"]["shadow"]) or strategy == "zero2"
            # load_state_dict scatters a full fp32 copy back to the owners: a perturbed copy round-trips
            sd2 = {k: v + 0.5 for k, v in sd.items()}
            model.load_state_dict(sd2)
            again = model.state_dict()
            assert all(torch.equal(again[k], sd2[k]) for k in sd2)
            assert torch.equal(f.pview("l0.bias"), sd2["l0.bias"])

        # ZeRO-1 checkpoint: 2 steps, save, fresh engine (different init), load, third step == uninterrupted
        model = Toy()
        eng = TrainEngine(model, ToyAdam(model.flat), None, max_grad_norm=0.05, gradient_accumulation_steps=2, strategy="zero1")
        run(eng, range(2))
        eng.save_checkpoint(tmpdir)
        model2 = Toy(seed=5)
        eng2 = TrainEngine(model2, ToyAdam(model2.flat), None, max_grad_norm=0.05, gradient_accumulation_steps=2, strategy="zero1")
        eng2.load_checkpoint(tmpdir)
        assert eng2.micro == 4 and eng2.optimizer.t == 2
        run(eng2, range(2, 3))
        model2.state_dict()
        assert torch.equal(model2.flat.master, finals["zero1"]["master"])
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback

        q.put((rank, "".join(traceback.format_exception(e))[-1500:]))
    finally:
        dist.destroy_process_group()


def test_train_engine_zero1_equals_ddp_and_resumes_world2(tmp_path):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_engine_worker, args=(r, 2, port, q, str(tmp_path / "ckpt"))) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


# ---------------------------------------------------------------------------------------------------------------
# fp16 overflow handling across ranks: a non-finite gradient on ONE rank must skip the optimizer step on EVERY rank (the flag comes
# from the all-reduced gradients under DDP and from the all-reduced sum of squares under ZeRO), leave parameters and moments
# untouched, clear the accumulators (the ZeRO-2 shard too), back the loss scale off once, and keep the scheduler where it was.
# clip_coef / loss_scale_update are torch statements of adam.cu's clip_coef_kernel / loss_scale_update_kernel.
# ---------------------------------------------------------------------------------------------------------------
def _clip_coef_ref(sumsq, max_norm, loss_scale=None, found_inf=None):
    s = float(loss_scale) if loss_scale is not None else 1.0
    if found_inf is not None:
        found_inf.fill_(0 if bool(torch.isfinite(sumsq)) else 1)
    norm = sumsq.sqrt() / s
    coef = torch.clamp(max_norm / (norm + 1e-6), max=1.0) if max_norm and max_norm > 0 else torch.ones(())
    return norm, coef / s


def _loss_scale_update_ref(scale, tracker, hyst_left, found_inf, growth_factor, backoff_factor, growth_interval, min_scale, hysteresis):
    if int(found_inf):
        left = int(hyst_left) - 1
        if hysteresis <= 1 or left <= 0:
            scale.fill_(max(float(scale) * backoff_factor, min_scale))
            left = hysteresis if hysteresis <= 1 else 1
        hyst_left.fill_(left)
        tracker.zero_()
    else:
        t = int(tracker) + 1
        if t >= growth_interval:
            scale.mul_(growth_factor)
            tracker.zero_()
            hyst_left.fill_(hysteresis)
        else:
            tracker.fill_(t)


def _overflow_worker(rank, world, port, q, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import multimodal_llm_pretraining_b200.kernels as K
        from multimodal_llm_pretraining_b200.engine import LossScaler, TrainEngine

        K.sumsq_ = lambda x, out: out.add_((x.double() ** 2).sum().float())
        K.clip_coef = _clip_coef_ref
        K.loss_scale_update = _loss_scale_update_ref
        Toy, ToyAdam = _toy_classes()
        g = torch.Generator().manual_seed(100)
        xs = torch.randint(0, 32, (4, 2, world, 16), generator=g)
        ys = torch.randn(4, 2, world, 16, 4, generator=g)
        ys[1, 0, 1, 3, 2] = float("inf")  # step 1, first micro-batch, rank 1 only

        class Sched:
            n = 0

            def step(self):
                self.n += 1

            def state_dict(self):
                return {"n": self.n}

            def load_state_dict(self, sd):
                self.n = sd["n"]

        finals = {}
        for strategy in ("ddp", "zero1", "zero2"):
            model, sched = Toy(), Sched()
            scaler = LossScaler("cpu", kind="torch", init_scale=1024.0, growth_interval=2)
            eng = TrainEngine(model, ToyAdam(model.flat), sched, max_grad_norm=0.05, gradient_accumulation_steps=2, strategy=strategy,
                              loss_scaler=scaler)
            assert model.loss_scale is scaler.scale
            f, done = model.flat, []
            for s in range(4):
                for m in range(2):
                    eng.manual_training_step({"input_ids": xs[s, m, rank], "labels": ys[s, m, rank]})
                if s == 1:
                    model.state_dict()  # consolidates the fp32 master under ZeRO-1/2
                    before = (f.master.clone(), f.shadow.clone(), eng.optimizer._m.clone(), eng.optimizer._v.clone())
                ok = eng.manual_optimization_step()
                done.append(ok)
                votes = torch.tensor([int(ok)])
                dist.all_reduce(votes)
                assert int(votes) in (0, world), "ranks disagree about skipping step %d" % s
                if s == 1:
                    assert not ok and float(scaler.scale) == 512.0 and scaler.skipped_steps == 1 and int(scaler.growth_tracker) == 0
                    assert eng.optimizer.t == 1 and sched.n == 1
                    model.state_dict()
                    after = (f.master, f.shadow, eng.optimizer._m, eng.optimizer._v)
                    assert all(torch.equal(a, b) for a, b in zip(before, after)), "a skipped step must not move parameters or moments"
                    acc = eng._gshard if strategy == "zero2" else f.grad
                    assert float(acc.abs().max()) == 0.0 and f.pending_grad_scale is None
                    if strategy == "zero1":  # checkpoint carries the backed-off scale
                        eng.save_checkpoint(tmpdir)
                        model2 = Toy(seed=3)
                        eng2 = TrainEngine(model2, ToyAdam(model2.flat), Sched(), max_grad_norm=0.05, gradient_accumulation_steps=2,
                                           strategy=strategy, loss_scaler=LossScaler("cpu", kind="torch", init_scale=1024.0, growth_interval=2))
                        eng2.load_checkpoint(tmpdir)
                        assert float(eng2.loss_scaler.scale) == 512.0 and eng2.loss_scaler.skipped_steps == 1
            assert done == [True, False, True, True]
            assert float(scaler.scale) == 1024.0 and eng.optimizer.t == 3 and sched.n == 3  # two clean steps after the back-off: grown again
            model.state_dict()
            assert bool(torch.isfinite(f.master).all())
            finals[strategy] = f.master.clone()
        # power-of-two scales are exact in fp32, so the scaled ZeRO-1 run is the scaled DDP run bit for bit
        assert torch.equal(finals["zero1"], finals["ddp"])
        moved = (finals["ddp"] - Toy().flat.master).norm()
        assert ((finals["zero2"] - finals["ddp"]).norm() / moved).item() < 1e-5
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback

        q.put((rank, "".join(traceback.format_exception(e))[-1500:]))
    finally:
        dist.destroy_process_group()


def test_fp16_overflow_on_one_rank_skips_the_step_everywhere_world2(tmp_path):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_overflow_worker, args=(r, 2, port, q, str(tmp_path / "ckpt"))) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_zero3_refuses_modules_without_backward_weight_hooks():
    """ZeRO-3 gathers weights per bucket on the module's announcement; a module that does not announce its backward reads would read
    released buffers, so the engine refuses it up front."""
    from multimodal_llm_pretraining_b200.engine import TrainEngine

    Toy, ToyAdam = _toy_classes()
    model = Toy()
    with pytest.raises(NotImplementedError, match="zero3"):
        TrainEngine(model, ToyAdam(model.flat), None, strategy="zero3")
