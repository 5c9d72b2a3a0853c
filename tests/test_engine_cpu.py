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
