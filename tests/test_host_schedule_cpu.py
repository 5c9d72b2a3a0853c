"""CPU tests of the HOST side of the hot path with the REAL modules: B200GPTNeoXForCausalLM / B200RobertaForMaskedLM's hand-scheduled
forward and backward, the flat store's routing, B200Adam's chunk tables and the TrainEngine strategies run here on CPU tensors with the
C-ABI wrappers replaced by torch statements of their contracts (tests/cpu_kernels.py, test infrastructure). What is checked is
everything ABOVE the kernels: which tensor feeds which GEMM in which layout, where every gradient is accumulated, what the
optimizer touches, what the engine exchanges — against the HF golden fixtures and against each other across strategies (gloo,
world_size 2). The kernels themselves are checked on the GPU (`-m gpu`); the product's own CPU guards stay in place (the test
subclasses override `_require_cuda`, nothing else)."""
import math
import os
import sys
from pathlib import Path
from types import SimpleNamespace

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import cpu_kernels  # noqa: E402
from multimodal_llm_pretraining_b200.modeling_gpt_neox import B200GPTNeoXForCausalLM  # noqa: E402
from multimodal_llm_pretraining_b200.modeling_roberta import B200RobertaForMaskedLM  # noqa: E402
from multimodal_llm_pretraining_b200.optim import B200Adam, B200AdamW  # noqa: E402

GOLD = ROOT / "tests" / "golden" / "neox_tiny.pt"
GOLD_ROBERTA = ROOT / "tests" / "golden" / "roberta_tiny.pt"


class CpuNeoX(B200GPTNeoXForCausalLM):
    def _require_cuda(self) -> None:  # the kernels are replaced by cpu_kernels in these tests
        pass


class CpuRoberta(B200RobertaForMaskedLM):
    def _require_cuda(self) -> None:
        pass


class CpuAdam(B200Adam):
    def _require_cuda(self) -> None:
        pass


class CpuAdamW(B200AdamW):
    def _require_cuda(self) -> None:
        pass


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


@pytest.fixture()
def K(monkeypatch):
    return cpu_kernels.install(monkeypatch)


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, map_location="cpu", weights_only=False)


def _neox(gold):
    m = CpuNeoX(SimpleNamespace(**gold["cfg"]))
    m.load_hf_state_dict(gold["state_dict"])
    return m.train()


def test_product_guards_stay_in_place():
    """Without the test subclass the real classes refuse to run on CPU (no fallback in the product)."""
    cfg = dict(vocab_size=512, hidden_size=128, num_hidden_layers=1, num_attention_heads=2, intermediate_size=512, rotary_pct=0.25,
               rotary_emb_base=10000, layer_norm_eps=1e-5)
    m = B200GPTNeoXForCausalLM(SimpleNamespace(**cfg))
    ids = torch.randint(0, 512, (1, 9))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(input_ids=ids, labels=ids)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        B200Adam(m.parameters()).step()


def test_neox_schedule_loss_and_every_gradient_vs_hf_golden(K, gold):
    m = _neox(gold)
    ids = gold["batches"][0]
    loss = m(input_ids=ids, labels=ids).loss
    assert abs(loss.item() - gold["loss0"]) <= 2e-3 * gold["loss0"]
    loss.backward()
    grads = {n: p.grad for n, p in m.named_parameters()}
    nh = gold["cfg"]["num_attention_heads"]
    for k, g in gold["grads"].items():
        a, b = grads[k], g
        if k.endswith("query_key_value.bias"):  # the key third has an analytically zero gradient (tests/test_model_gpu.py)
            a, b = a.view(nh, 3, -1)[:, [0, 2]], g.view(nh, 3, -1)[:, [0, 2]]
        assert rel(a, b) <= 2e-2, (k, rel(a, b))
    for k, n in gold["grad_norms"].items():  # every parameter's gradient norm (the fixture keeps full tensors for a subset)
        assert abs(grads[k].norm().item() - n) <= 5e-2 * n + 1e-6, k


def test_neox_three_fused_adam_steps_vs_hf_golden(K, gold):
    m = _neox(gold)
    decay = [p for n, p in m.named_parameters() if p.dim() >= 2]
    no_decay = [p for n, p in m.named_parameters() if p.dim() < 2]
    opt = CpuAdam([{"params": decay, "weight_decay": 0.0}, {"params": no_decay, "weight_decay": 0.0}], lr=6e-4, betas=(0.9, 0.95), eps=1e-8)
    losses, norms = [], []
    for ids in gold["batches"]:
        loss = m(input_ids=ids, labels=ids).loss
        loss.backward()
        norms.append(m.clip_grad_norm_(1.0).item())
        opt.step()
        m.zero_grad()
        losses.append(loss.item())
    for a, b in zip(losses, gold["losses"]):
        assert abs(a - b) <= 1e-2 * b, (losses, gold["losses"])
    for a, b in zip(norms, gold["clip_norms"]):
        assert abs(a - b) <= 5e-2 * b, (norms, gold["clip_norms"])
    sd = m.state_dict()
    nh = gold["cfg"]["num_attention_heads"]
    for k, v in gold["params_after3"].items():
        upd_ref, upd_got = v - gold["state_dict"][k], sd[k] - gold["state_dict"][k]
        if k.endswith("query_key_value.bias"):
            upd_ref, upd_got = upd_ref.view(nh, 3, -1)[:, [0, 2]], upd_got.view(nh, 3, -1)[:, [0, 2]]
        assert rel(upd_got, upd_ref) <= 0.1, (k, rel(upd_got, upd_ref))
    # the 16-bit compute copy is what the next forward reads: it must be the rounded master everywhere
    f = m.flat
    assert torch.equal(f.shadow, f.master.to(f.shadow.dtype))


def test_neox_checkpointing_and_accumulation_are_exact_on_the_host_side(K, gold):
    """Recomputing a layer in backward replays the same statements: bit-identical gradients. Two micro-batches accumulate into the
    same buffer: the sum of the two single-batch gradients (fp32 accumulation order is fixed on CPU)."""
    ids0, ids1 = gold["batches"][0], gold["batches"][1]
    m = _neox(gold)
    m(input_ids=ids0, labels=ids0).loss.backward()
    g0 = m.flat.grad.clone()
    m.zero_grad()
    m.gradient_checkpointing_enable()
    m(input_ids=ids0, labels=ids0).loss.backward()
    assert torch.equal(m.flat.grad, g0)
    m.gradient_checkpointing_disable()
    m.zero_grad()
    m(input_ids=ids1, labels=ids1).loss.backward()
    g1 = m.flat.grad.clone()
    m.zero_grad()
    m(input_ids=ids0, labels=ids0).loss.backward()
    m(input_ids=ids1, labels=ids1).loss.backward()
    assert rel(m.flat.grad, g0 + g1) <= 1e-6
    # the buckets the engine exchanges are fired once each, in backward order, and tile the whole store
    fired = []
    m.zero_grad()
    m.grad_ready_hook = lambda s, e: fired.append((s, e))
    m(input_ids=ids0, labels=ids0).loss.backward()
    assert fired == m.comm_buckets()
    assert sorted(fired)[0][0] == 0 and all(a[1] == b[0] for a, b in zip(sorted(fired), sorted(fired)[1:]))


def test_neox_eval_and_logits_paths(K, gold):
    m = _neox(gold).eval()
    ids = gold["batches"][0]
    out = m(input_ids=ids, labels=ids)
    assert abs(out.loss.item() - gold["loss0"]) <= 2e-3 * gold["loss0"]
    B, S = ids.shape
    assert out.logits.shape == (B, S - 1, gold["cfg"]["vocab_size"])
    full = m(input_ids=ids).logits
    assert full.shape == (B, S, gold["cfg"]["vocab_size"])
    assert rel(full[:, :-1], out.logits) <= 1e-6  # causal: dropping the last token does not change the others
    if "logits0" in gold:
        assert rel(full, gold["logits0"]) <= 2e-2


# ------------------------------------------------------------------------------------------------------------- RoBERTa
@pytest.fixture(scope="module")
def gold_r():
    return torch.load(GOLD_ROBERTA, map_location="cpu", weights_only=False)


def _roberta(gold_r):
    m = CpuRoberta(SimpleNamespace(**gold_r["cfg"]))
    m.load_hf_state_dict(gold_r["state_dict"])
    return m.train()


def test_roberta_schedule_logits_loss_and_every_gradient_vs_hf_golden(K, gold_r):
    m = _roberta(gold_r)
    ids = gold_r["batches"][0]
    m.eval()
    assert rel(m(input_ids=ids, labels=ids)["logits"], gold_r["logits0"]) <= 2e-2
    m.train()
    loss = m(input_ids=ids, labels=ids)["loss"]
    assert abs(loss.item() - gold_r["loss0"].item()) <= 5e-3 * gold_r["loss0"].item()
    loss.backward()
    grads = {n: p.grad for n, p in m.named_parameters()}
    for k, g in gold_r["grads0"].items():
        if k.endswith("attention.self.key.bias"):  # analytically zero (tests/test_roberta_gpu.py)
            qn = gold_r["grads0"][k.replace("key.bias", "query.bias")].norm().item()
            assert grads[k].float().norm().item() <= 2e-2 * qn, k
            continue
        assert rel(grads[k], g) <= 5e-2, (k, rel(grads[k], g))


def test_roberta_three_fused_adamw_steps_vs_hf_golden(K, gold_r):
    m = _roberta(gold_r)
    opt = CpuAdam(m.parameters(), lr=4e-4, betas=(0.9, 0.98), weight_decay=0.0)
    losses = []
    for b in gold_r["batches"]:
        loss = m(input_ids=b, labels=b)["loss"]
        loss.backward()
        opt.step()
        m.zero_grad()
        losses.append(loss.item())
    for a, b in zip(losses, gold_r["losses_3steps"].tolist()):
        assert abs(a - b) <= 1e-2 * b, (losses, gold_r["losses_3steps"].tolist())
    sd = m.state_dict()
    for k in ("roberta.encoder.layer.1.intermediate.dense.weight", "roberta.embeddings.word_embeddings.weight", "lm_head.bias"):
        upd, ref_upd = sd[k] - gold_r["state_dict"][k], gold_r["state_dict_after3"][k] - gold_r["state_dict"][k]
        assert rel(upd, ref_upd) <= 0.25, (k, rel(upd, ref_upd))
    f = m.flat  # the vocabulary padding behind the embedding matrix / decoder bias stays exactly zero
    assert torch.all(f.view_alloc(f.master, "roberta.embeddings.word_embeddings.weight")[m.V:] == 0)
    assert torch.all(f.view_alloc(f.master, "lm_head.bias")[m.V:] == 0)


def test_roberta_checkpointing_is_exact_and_buckets_fire_in_backward_order(K, gold_r):
    ids = gold_r["batches"][0]
    m = _roberta(gold_r)
    m(input_ids=ids, labels=ids)["loss"].backward()
    g0 = m.flat.grad.clone()
    m.zero_grad()
    m.gradient_checkpointing_enable()
    fired = []
    m.grad_ready_hook = lambda s, e: fired.append((s, e))
    m(input_ids=ids, labels=ids)["loss"].backward()
    assert torch.equal(m.flat.grad, g0)
    assert fired == m.comm_buckets()
    srt = sorted(fired)
    assert srt[0][0] == 0 and all(a[1] == b[0] for a, b in zip(srt, srt[1:]))


# ------------------------------------------------------------------------------------------------------------- TrainEngine, world 2
# The CPU twin of scripts/dev/dp_check.py (which runs on 2 / 8 GPUs): the REAL NeoX and RoBERTa modules through the REAL TrainEngine over
# gloo — DDP, ZeRO-1, ZeRO-2, the sharded fp32 master, activation checkpointing, fp16 + loss scaling — compared with a single-process
# run over the same global batch and with each other.
TINY = dict(vocab_size=256, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=512, rotary_pct=0.25,
            rotary_emb_base=10000, layer_norm_eps=1e-5)
TINY_R = dict(vocab_size=301, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, max_position_embeddings=66,
              type_vocab_size=1, pad_token_id=1, layer_norm_eps=1e-5, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
              hidden_act="gelu", initializer_range=0.02)


def _build(kind="neox", seed=0):
    m = CpuNeoX(SimpleNamespace(**TINY)) if kind == "neox" else CpuRoberta(SimpleNamespace(**TINY_R))
    m.reset_parameters(torch.Generator().manual_seed(seed))
    with torch.no_grad():
        g = torch.Generator().manual_seed(seed + 1)
        for n, p in m.named_parameters():
            if n.endswith(".bias"):
                p.normal_(0, 0.02, generator=g)
    m.flat.sync_shadow(force=True)
    return m.train()


def _noise_mask(m):
    """False on the elements whose gradient is analytically zero (key biases): Adam turns their rounding noise into +-lr steps."""
    keep = torch.ones(m.flat.numel, dtype=torch.bool)
    for n, p in m.named_parameters():
        _, off, cnt = p._b200_flat
        if n.endswith("query_key_value.bias"):
            keep[off:off + cnt].view(m.nh, 3, -1)[:, 1] = False
        if n.endswith("attention.self.key.bias"):
            keep[off:off + cnt] = False
    return keep


class _FakeEvent:
    def __init__(self, *a, **k):
        pass

    def record(self, *a):
        pass


class _FakeStream:
    def wait_event(self, ev):
        assert isinstance(ev, _FakeEvent)

    def wait_stream(self, st):
        assert isinstance(st, _FakeStream)


def _fake_side_stream(eng):
    """Drive the engine's OVERLAP branches (side stream + events: the default on a GPU) on CPU tensors: streams and events become no-ops,
    so every collective runs at the point where it is enqueued — program order, which is one of the orders the real streams allow. What
    this checks is the bookkeeping of those branches (ring buffers, in-flight gathers, per-bucket events); the ordering itself is a
    property of the CUDA streams and is checked on hardware by scripts/dev/dp_check.py."""
    import contextlib

    torch.cuda.Event = _FakeEvent
    torch.cuda.current_stream = lambda *a, **k: _FakeStream()
    torch.cuda.stream = lambda st: contextlib.nullcontext()
    eng.comm_stream = _FakeStream()
    assert eng.overlap and eng.overlap_plan is not None


def _dp_worker(rank, world, port, q, tmpdir, kind):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cpu_kernels.install()
        from multimodal_llm_pretraining_b200.engine import LossScaler, TrainEngine

        steps, ga, B = 2, 2, 2
        S = 17 if kind == "neox" else 16
        V = TINY["vocab_size"] if kind == "neox" else TINY_R["vocab_size"]
        data = torch.randint(2, V, (steps, ga, world, B, S), generator=torch.Generator().manual_seed(7))
        Opt = CpuAdam if kind == "neox" else CpuAdamW

        def train(strategy, **kw):
            dtype = kw.pop("dtype", None)
            ckpt = kw.pop("ckpt", False)
            model = _build(kind)
            if dtype is not None:
                model.set_compute_dtype(dtype)
                kw["loss_scaler"] = LossScaler("cpu", kind="torch", init_scale=256.0)
            if ckpt:
                model.gradient_checkpointing_enable()
            opt = Opt(model.parameters(), lr=1e-3, betas=(0.9, 0.95), weight_decay=0.0)
            side = kw.pop("side", False)
            if strategy == "none":  # one process, the whole global batch: micro-batches of every rank in turn
                eng = TrainEngine(model, opt, None, max_grad_norm=0.5, gradient_accumulation_steps=ga * world, strategy="none", **kw)
                for s in range(steps):
                    for m in range(ga):
                        for r in range(world):
                            eng.manual_training_step({"input_ids": data[s, m, r], "labels": data[s, m, r]})
                    assert eng.manual_optimization_step()
            else:
                eng = TrainEngine(model, opt, None, max_grad_norm=0.5, gradient_accumulation_steps=ga, strategy=strategy, **kw)
                if side:
                    _fake_side_stream(eng)
                for s in range(steps):
                    for m in range(ga):
                        eng.manual_training_step({"input_ids": data[s, m, rank], "labels": data[s, m, rank]})
                    assert eng.manual_optimization_step()
            assert float(eng.last_grad_norm) > 0.5, "clipping must be active in this test"
            full = model.flat.materialize_master().clone()
            return model, eng, opt, full

        init = _build(kind).flat.master.clone()
        keep = _noise_mask(_build(kind))

        def diff(a, b):
            ua, ub = (a - init)[keep], (b - init)[keep]
            return ((ua - ub).norm() / ub.norm()).item()

        _, _, _, ref = train("none")
        assert (ref - init).abs().max() > 1e-4
        finals, norms = {}, {}
        for strategy in ("ddp", "zero1", "zero2"):
            model, eng, opt, full = train(strategy)
            finals[strategy] = full
            norms[strategy] = float(eng.last_grad_norm)
            f = model.flat
            assert torch.equal(f.shadow, full.to(f.shadow.dtype)), "16-bit compute copy != rounded fp32 master"
            lst = [torch.empty_like(full) for _ in range(world)]
            dist.all_gather(lst, full)
            assert all(torch.equal(lst[0], x) for x in lst), "ranks hold different parameters"
            # same mean gradient as the single-process run, summed in a different order, through bf16 activations
            assert diff(full, ref) < 2e-2, (strategy, diff(full, ref))
            if strategy != "ddp":
                assert f.numel - 8 * 64 <= opt._m.numel() * world <= f.numel  # the store's tail padding belongs to no bucket
            if strategy == "zero2":
                assert f.grad is None and eng._gshard.numel() == opt._m.numel() and float(eng._gshard.abs().max()) == 0.0
        # two-term sums are order independent: ZeRO-1's reduce-scatter and DDP's all-reduce give the same gradients; the norm is reduced in
        # a different order (owned chunks + scalar all-reduce), so the clip coefficient may differ in the last ulp
        assert diff(finals["zero1"], finals["ddp"]) < 1e-5
        assert diff(finals["zero2"], finals["ddp"]) < 1e-2
        # sharded fp32 master: the same arithmetic in a different place -> bit for bit, incl. a checkpoint round trip and the next step
        for strategy in ("zero1", "zero2"):
            model, eng, opt, full = train(strategy, shard_master=True)
            assert model.flat.master is None and opt._p32.numel() == opt._m.numel()
            assert torch.equal(full, finals[strategy]), strategy
            eng.save_checkpoint(os.path.join(tmpdir, strategy))
            m2 = _build(kind, seed=5)
            o2 = Opt(m2.parameters(), lr=1e-3, betas=(0.9, 0.95), weight_decay=0.0)
            e2 = TrainEngine(m2, o2, None, max_grad_norm=0.5, gradient_accumulation_steps=ga, strategy=strategy, shard_master=True)
            e2.load_checkpoint(os.path.join(tmpdir, strategy))
            assert torch.equal(m2.flat.materialize_master(), full) and torch.equal(m2.flat.shadow, model.flat.shadow)
            for e_ in (eng, e2):
                for m in range(ga):
                    e_.manual_training_step({"input_ids": data[0, m, rank], "labels": data[0, m, rank]})
                e_.manual_optimization_step()
            assert torch.equal(m2.flat.materialize_master(), model.flat.materialize_master())
        # ZeRO-3: the 16-bit weights are sharded too and gathered bucket by bucket through the module's parameter hooks. Same arithmetic
        # as ZeRO-2 with a sharded master -> bit for bit; the parameters are empty placeholders (DeepSpeed stage-3 style)
        if True:
            for ckpt in (False, True):
                model, eng, opt, full = train("zero3", ckpt=ckpt)
                f = model.flat
                assert f.shadow is None and f.master is None and f.grad is None
                assert eng._w16.numel() == opt._m.numel() == opt._p32.numel() == eng._gshard.numel()
                assert all(p.numel() == 0 for p in model.parameters()) and model.num_parameters() == _build(kind).num_parameters()
                all_sizes = [e - s_ for s_, e in eng.plan.buckets]  # embedding, layers (one ring of two), head — independent of the depth
                assert eng.zero3_transient_bytes() == 2 * sum(n * (2 if all_sizes.count(n) > 1 else 1) for n in set(all_sizes))
                assert torch.equal(full, finals["zero2"]), ("zero3", ckpt)
            sd = model.state_dict()
            assert set(sd) == set(_build(kind).state_dict()) and all(v.dtype == torch.float32 for v in sd.values())
            eng.save_checkpoint(os.path.join(tmpdir, "zero3"))
            m2 = _build(kind, seed=5)
            o2 = Opt(m2.parameters(), lr=1e-3, betas=(0.9, 0.95), weight_decay=0.0)
            e2 = TrainEngine(m2, o2, None, max_grad_norm=0.5, gradient_accumulation_steps=ga, strategy="zero3")
            e2.load_checkpoint(os.path.join(tmpdir, "zero3"))
            assert torch.equal(m2.flat.materialize_master(), full) and torch.equal(e2._w16, eng._w16)
            for e_ in (eng, e2):
                for m in range(ga):
                    e_.manual_training_step({"input_ids": data[0, m, rank], "labels": data[0, m, rank]})
                e_.manual_optimization_step()
            assert torch.equal(m2.flat.materialize_master(), model.flat.materialize_master())
            model.eval()  # the forward-only paths announce their buckets too
            with torch.no_grad():
                ev = model(input_ids=data[0, 0, rank], labels=data[0, 0, rank])
            assert bool(torch.isfinite(ev.loss)) and ev.logits.shape[-1] == V
            # a weight read outside the hooks is an error, not a silent read of a stale buffer
            eng._invalidate_weights()
            with pytest.raises(RuntimeError, match="not resident"):
                model._w("embed_out.weight" if kind == "neox" else "lm_head.dense.weight")
        # the overlap branches (side stream, events, in-flight gathers, ring reuse) with no-op streams: same results. LAST in this
        # worker: torch.cuda.Event / stream / current_stream stay patched afterwards
        tail = [("ddp", {}), ("zero1", {}), ("zero2", {}), ("zero1", {"shard_master": True})] + [("zero3", {}), ("zero3", {"ckpt": True})]
        # activation checkpointing recomputes the same statements: bit-identical to the plain run of the same strategy
        _, _, _, full = train("zero1", ckpt=True)
        assert torch.equal(full, finals["zero1"])
        # fp16 + loss scaling: a different rounding of the same computation (power-of-two scale: exact un-scaling)
        model, eng, _, full = train("zero1", dtype=torch.float16)
        assert model.flat.shadow.dtype == torch.float16 and eng.loss_scaler.skipped_steps == 0
        # Adam's update is nearly invariant to the gradient scale, so the un-scaling is checked on the reported gradient norm ...
        assert abs(float(eng.last_grad_norm) - norms["zero1"]) < 2e-2 * norms["zero1"], (float(eng.last_grad_norm), norms["zero1"])
        # ... and the update against the bf16 run of the same strategy: two roundings (8 vs 11 significant bits) of one computation
        assert diff(full, finals["zero1"]) < 0.15, diff(full, finals["zero1"])
        fp16_zero1 = full
        model, eng, _, full = train("zero3", dtype=torch.float16)   # ZeRO-3 in fp16: the 16-bit shard is fp16, same arithmetic as ZeRO-2
        assert eng._w16.dtype == torch.float16 and eng.loss_scaler.skipped_steps == 0
        assert diff(full, fp16_zero1) < 1e-2, diff(full, fp16_zero1)
        for strategy, kw in tail:
            _, eng, _, full = train(strategy, side=True, **kw)
            assert torch.equal(full, finals["zero2" if strategy == "zero3" else strategy]), ("side stream", strategy, kw)
            if strategy in ("zero1",) and not kw:
                assert eng._param_events, "the per-bucket gather events of the side-stream branch were not recorded"
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback

        q.put((rank, "".join(traceback.format_exception(e))[-2500:]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["neox", "roberta"])
def test_real_modules_through_the_engine_world2(tmp_path, kind):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 35500 + os.getpid() % 2000 + (0 if kind == "neox" else 1)
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q, str(tmp_path), kind)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_torch_optimizer_interop_recasts_the_compute_copy(K, gold):
    """A foreign optimizer (the reference's naive path hands out torch.optim.Adam) edits the fp32 master through torch: the version
    counters move, the next forward re-casts the 16-bit copy; zero_grad(set_to_none=True) drops the gradient views and the module
    re-attaches them."""
    m = _neox(gold)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    ids = gold["batches"][0]
    l0 = m(input_ids=ids, labels=ids).loss
    l0.backward()
    before = m.flat.shadow.clone()
    opt.step()
    opt.zero_grad(set_to_none=True)
    assert next(m.parameters()).grad is None
    assert torch.equal(m.flat.shadow, before)                       # not yet re-cast ...
    l1 = m(input_ids=ids, labels=ids).loss
    assert torch.equal(m.flat.shadow, m.flat.master.to(m.flat.shadow.dtype)) and not torch.equal(m.flat.shadow, before)  # ... now it is
    assert l1.item() < l0.item()
    l1.backward()
    p = next(m.parameters())
    assert p.grad is not None and p.grad.data_ptr() == m.flat.view(m.flat.grad, "gpt_neox.embed_in.weight").data_ptr()
    assert float(m.flat.grad.abs().max()) > 0


def test_roberta_dropout_seed_plumbing(K, gold_r):
    """Both dropouts on (the roberta-large configuration). The masks are functions of (step seed, site, element): backward and
    activation recomputation must regenerate the forward's masks, every training forward must draw new ones, eval must draw none, and
    the analytic gradient must agree with a finite difference of the SAME masked function (a backward that used other masks would not)."""
    cfg = dict(gold_r["cfg"], hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1)
    ids = gold_r["batches"][0]

    def build():
        m = CpuRoberta(SimpleNamespace(**cfg))
        m.load_hf_state_dict(gold_r["state_dict"])
        return m.train()

    m = build()
    l1 = m(input_ids=ids, labels=ids)["loss"]
    l1.backward()
    g_plain = m.flat.grad.clone()
    l2 = m(input_ids=ids, labels=ids)["loss"]
    assert m._step_seed == 2 and l1.item() != l2.item()            # same batch, same weights, new masks
    m.eval()
    e1, e2 = m(input_ids=ids, labels=ids)["loss"].item(), m(input_ids=ids, labels=ids)["loss"].item()
    assert e1 == e2 and m._step_seed == 2                             # no dropout, no seed consumed
    # recomputation in backward regenerates the same masks: bit-identical gradients for the same step seed
    mc = build()
    mc.gradient_checkpointing_enable()
    lc = mc(input_ids=ids, labels=ids)["loss"]
    assert lc.item() == l1.item()
    lc.backward()
    assert torch.equal(mc.flat.grad, g_plain)
    # a forward of ANOTHER step between forward and backward must not change the masks the backward uses
    mi = build()
    la = mi(input_ids=ids, labels=ids)["loss"]
    mi(input_ids=ids, labels=ids)["loss"]                            # bumps the step seed, graph discarded
    la.backward()
    assert torch.equal(mi.flat.grad, g_plain)
    # directional finite difference of the masked loss (step seed pinned) against the analytic gradient
    name = "roberta.encoder.layer.0.output.dense.bias"                 # fp32 parameter read directly by the kernels: no 16-bit rounding of the step
    mf = build()
    p = dict(mf.named_parameters())[name]
    d = torch.randn(p.shape, generator=torch.Generator().manual_seed(0))
    d = d / d.norm()

    def loss_at(eps):
        mf._step_seed = 0
        with torch.no_grad():
            p.add_(eps * d)
        out = mf(input_ids=ids, labels=ids)["loss"].item()
        with torch.no_grad():
            p.sub_(eps * d)
        return out

    fd = (loss_at(0.05) - loss_at(-0.05)) / 0.1
    an = float((mf.flat.view(g_plain, name) * d).sum())
    assert abs(fd - an) <= 0.15 * abs(an) + 2e-3, (fd, an)


def test_reference_surface_to_engine_on_cpu(K, monkeypatch):
    """TrainingConfig -> TrainingClass.build_trainer -> ManualTrainer (optimizer class hand-off, HF's two parameter groups, scheduler from
    the TrainingArguments dict) -> the reference's two harness calls, on a tiny module."""
    from multimodal_llm_pretraining_b200.benchmarking.utils import ManualTrainer
    from multimodal_llm_pretraining_b200.config import TrainingConfig

    monkeypatch.setattr(B200Adam, "_require_cuda", lambda self: None)
    config = TrainingConfig(1, 1, "b200", "pythia-70m", free_lunch=True)
    tc = config.training_class(num_training_steps=4, micro_batch_size=2, gradient_accumulation_steps=2, bf16=True, fp16=False)
    assert tc.is_valid() and tc.runs_on_b200_engine()
    mc = config.model_class()
    assert mc.optimizer is torch.optim.Adam                       # the registry hands out the reference's own class ...
    model = _build("neox")
    ds = mc.load_dummy_dataset(num_samples=16, seed=0)
    ds.input_ids %= TINY["vocab_size"]
    ds.labels %= TINY["vocab_size"]
    ds.input_ids, ds.labels = ds.input_ids[:, :33], ds.labels[:, :33]
    trainer = tc.build_trainer(model, ds, hf_trainer_kwargs_overrides={"device": torch.device("cpu")})
    assert isinstance(trainer, ManualTrainer) and type(trainer.optimizer) is B200Adam   # ... the trainer swaps in the fused twin
    assert trainer.engine.strategy == "none" and trainer.engine.ga == 2
    groups = trainer.optimizer.param_groups
    assert len(groups) == 2 and all(p.dim() >= 2 for p in groups[0]["params"]) and all(p.dim() < 2 for p in groups[1]["params"])
    assert all(g["weight_decay"] == 0.0 for g in groups)            # TrainingArguments.weight_decay = 0 overrides the model class's kwarg
    assert groups[0]["betas"] == tuple(mc.optimizer_kwargs["betas"]) and groups[0]["eps"] == mc.optimizer_kwargs["eps"]
    it = iter(trainer.get_train_dataloader())
    lrs, losses = [], []
    before = model.flat.master.clone()
    for _ in range(3):
        for _ in range(2):
            losses.append(float(trainer.manual_training_step(trainer.model, next(it))))
        lrs.append(groups[0]["lr"])
        assert trainer.manual_optimization_step(trainer.model)
    assert trainer.optimizer._step == 3 and trainer.lr_scheduler.last_epoch == 3
    assert len(set(lrs)) > 1, "the schedule of the TrainingArguments dict must drive the fused optimizer's lr"
    assert all(math.isfinite(x) for x in losses) and not torch.equal(model.flat.master, before)
    assert float(model.flat.grad.abs().max()) == 0.0                 # zero_grad at the end of the optimization step
    # a stock HF-style module (no flat store) is refused with a message, not trained on a silent fallback
    with pytest.raises(NotImplementedError, match="B200 modules"):
        tc.build_trainer(torch.nn.Linear(4, 4), ds, hf_trainer_kwargs_overrides={"device": torch.device("cpu")})


def test_neox_opt_in_fused_bias_gradient_schedule(K, gold, monkeypatch):
    """B200_FUSED_BIAS_GRAD=1 moves the dense_h_to_4h bias gradient into the dGELU dgrad epilogue (measured slower, kept switchable):
    the schedule must still put the same gradient in the same place."""
    import multimodal_llm_pretraining_b200.modeling_gpt_neox as M

    ids = gold["batches"][0]
    m = _neox(gold)
    m(input_ids=ids, labels=ids).loss.backward()
    ref = m.flat.grad.clone()
    monkeypatch.setattr(M, "FUSED_BIAS_GRAD", True)
    m.zero_grad()
    m(input_ids=ids, labels=ids).loss.backward()
    for n, _ in m.named_parameters():
        a, b = m.flat.view(m.flat.grad, n), m.flat.view(ref, n)
        if n.endswith("dense_h_to_4h.bias"):
            assert rel(a, b) <= 1e-2, (n, rel(a, b))   # column sums of the fp32 accumulator instead of the rounded dh1
        else:
            assert torch.equal(a, b), n


@pytest.mark.parametrize("nh,h,pct", [(2, 256, 0.25), (1, 256, 0.25), (4, 320, 0.25), (2, 128, 1.0)])  # head_dim 128, 256, 80 (rot 20), full rotary
def test_neox_schedule_vs_fp32_oracle_other_head_shapes(K, nh, h, pct):
    """The same check against the fp32 CPU oracle for the other Pythia head shapes: rotary tables / packed-qkv views / strides are host
    logic that depends on (heads, head_dim, rotary fraction)."""
    from oracle import neox_oracle as O

    cfg = dict(vocab_size=512, hidden_size=h, num_hidden_layers=2, num_attention_heads=nh, intermediate_size=4 * h, rotary_pct=pct,
               rotary_emb_base=10000, layer_norm_eps=1e-5)
    m = CpuNeoX(SimpleNamespace(**cfg))
    m.reset_parameters(torch.Generator().manual_seed(nh))
    with torch.no_grad():
        g = torch.Generator().manual_seed(nh + 1)
        for n, p in m.named_parameters():
            if n.endswith(".bias"):
                p.normal_(0, 0.02, generator=g)
    m.train()
    ids = torch.randint(0, 512, (2, 65), generator=torch.Generator().manual_seed(5))
    P = {k: v.detach().clone() for k, v in m.state_dict().items()}
    ref_loss, ref_grads = O.neox_loss_and_grads(P, ids, ids, cfg)
    loss = m(input_ids=ids, labels=ids).loss
    loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 2e-3 * ref_loss.item()
    for n, p in m.named_parameters():
        a, b = p.grad, ref_grads[n]
        if n.endswith("query_key_value.bias"):
            a, b = a.view(nh, 3, -1)[:, [0, 2]], b.view(nh, 3, -1)[:, [0, 2]]
        assert rel(a, b) <= 2e-2, (n, rel(a, b))


def test_engine_single_process_equals_the_hand_written_loop(K, gold):
    """manual_training_step x ga + manual_optimization_step == the reference harness's statements written out (loss / ga backward,
    clip_grad_norm_, optimizer.step, scheduler.step, zero_grad; src/benchmarking/utils.py:61-80), with and without clipping."""
    from multimodal_llm_pretraining_b200.engine import TrainEngine
    from multimodal_llm_pretraining_b200.optim import get_scheduler

    b = gold["batches"]
    for max_norm in (1.0, 0.0):
        ma, mb = _neox(gold), _neox(gold)
        oa = CpuAdam(ma.parameters(), lr=1e-3, betas=(0.9, 0.95))
        ob = CpuAdam(mb.parameters(), lr=1e-3, betas=(0.9, 0.95))
        sa = get_scheduler("cosine_with_min_lr", oa, 1, 10, {"min_lr_rate": 0.1})
        sb = get_scheduler("cosine_with_min_lr", ob, 1, 10, {"min_lr_rate": 0.1})
        eng = TrainEngine(ma, oa, sa, max_grad_norm=max_norm, gradient_accumulation_steps=2, strategy="none")
        for step in range(2):
            for ids in (b[step], b[step + 1]):
                eng.manual_training_step({"input_ids": ids, "labels": ids})
                (mb(input_ids=ids, labels=ids).loss / 2).backward()
            assert eng.manual_optimization_step()
            if max_norm > 0:
                norm = mb.clip_grad_norm_(max_norm)
                assert float(norm) == float(eng.last_grad_norm)
            ob.step()
            sb.step()
            mb.zero_grad()
            assert oa.param_groups[0]["lr"] == ob.param_groups[0]["lr"]
        assert torch.equal(ma.flat.master, mb.flat.master) and torch.equal(ma.flat.shadow, mb.flat.shadow)
        assert eng.micro == 4 and oa._step == 2


def _dp_worker4(rank, world, port, q):
    """Four ranks: interior slices (ranks 1, 2) exist only beyond two ranks — packed-shard offsets, gather order, owned masks."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        cpu_kernels.install()
        from multimodal_llm_pretraining_b200.engine import TrainEngine

        steps, ga = 2, 2
        data = torch.randint(2, TINY["vocab_size"], (steps, ga, world, 2, 17), generator=torch.Generator().manual_seed(11))
        init = _build("neox").flat.master.clone()
        keep = _noise_mask(_build("neox"))

        def train(strategy, side=False, **kw):
            model = _build("neox")
            opt = CpuAdam(model.parameters(), lr=1e-3, betas=(0.9, 0.95), weight_decay=0.0)
            eng = TrainEngine(model, opt, None, max_grad_norm=0.5, gradient_accumulation_steps=ga, strategy=strategy, **kw)
            if side:
                _fake_side_stream(eng)
            for s in range(steps):
                for m in range(ga):
                    eng.manual_training_step({"input_ids": data[s, m, rank], "labels": data[s, m, rank]})
                assert eng.manual_optimization_step()
            full = model.flat.materialize_master().clone()
            lst = [torch.empty_like(full) for _ in range(world)]
            dist.all_gather(lst, full)
            assert all(torch.equal(lst[0], x) for x in lst), (strategy, "ranks hold different parameters")
            return eng, opt, full

        def diff(a, b):
            ua, ub = (a - init)[keep], (b - init)[keep]
            return ((ua - ub).norm() / ub.norm()).item()

        _, _, ddp = train("ddp")
        _, o1, z1 = train("zero1")
        assert o1._m.numel() * world <= init.numel() and diff(z1, ddp) < 1e-2       # gloo emulates both with the same all-reduce: ~0
        _, _, z1s = train("zero1", shard_master=True)
        assert torch.equal(z1s, z1)
        _, _, z2 = train("zero2")
        assert diff(z2, ddp) < 1e-2
        e3, o3, z3 = train("zero3")
        assert torch.equal(z3, z2) and e3._w16.numel() == o3._m.numel() == o3._p32.numel()
        _, _, z3s = train("zero3", side=True)                                      # LAST: torch.cuda stays patched
        assert torch.equal(z3s, z2)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback

        q.put((rank, "".join(traceback.format_exception(e))[-2500:]))
    finally:
        dist.destroy_process_group()


def test_real_module_through_the_engine_world4():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 37500 + os.getpid() % 2000
    procs = [ctx.Process(target=_dp_worker4, args=(r, 4, port, q)) for r in range(4)]
    for p in procs:
        p.start()
    res = [q.get(timeout=900) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(4)], res
