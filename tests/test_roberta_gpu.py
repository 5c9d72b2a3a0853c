"""Module-level parity on the B200: B200RobertaForMaskedLM + B200Adam (C-ABI kernels) against the golden vectors of the
real HF RobertaForMaskedLM (tests/golden/roberta_tiny.pt, dropout 0) and the fp32 CPU oracle; dropout statistics.
Tolerances vs fp32 references: bf16 operands => rel <= 2e-2 per tensor (5e-2 for whole-model parameter gradients)."""
import sys
from pathlib import Path
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from multimodal_llm_pretraining_b200 import kernels as K  # noqa: E402
from multimodal_llm_pretraining_b200.modeling_roberta import B200RobertaForMaskedLM  # noqa: E402
from multimodal_llm_pretraining_b200.optim import B200Adam  # noqa: E402
from oracle import roberta_oracle as R  # noqa: E402

GOLD = ROOT / "tests" / "golden" / "roberta_tiny.pt"


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, map_location="cpu", weights_only=False)


def build(gold, dev, **over):
    cfg = dict(gold["cfg"])
    cfg.update(over)
    m = B200RobertaForMaskedLM(SimpleNamespace(**cfg))
    m.load_hf_state_dict(gold["state_dict"])
    return m.to(dev).train()


def test_state_dict_keys_match_hf(gold, dev):
    m = build(gold, dev)
    hf_keys = set(gold["state_dict"]) | {"lm_head.decoder.weight", "lm_head.decoder.bias"}
    assert set(m.state_dict()) == hf_keys
    for k, v in gold["state_dict"].items():
        assert m.state_dict()[k].shape == v.shape, k
    assert m.state_dict()["lm_head.decoder.weight"].data_ptr() == m.state_dict()["roberta.embeddings.word_embeddings.weight"].data_ptr()
    # named_parameters lists every tensor once (the tied decoder adds nothing)
    assert sum(p.numel() for p in m.parameters()) == sum(v.numel() for v in gold["state_dict"].values())


def test_position_ids_bit_exact(dev):
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(0, 50, (7, 515), generator=g)
    ids[ids < 6] = 1  # plenty of pad tokens
    got = K.roberta_position_ids(ids.to(dev), 1).cpu()
    assert torch.equal(got, R.position_ids(ids, 1))


def test_loss_logits_and_grads_vs_hf_golden(gold, dev):
    m = build(gold, dev)
    ids = gold["batches"][0].to(dev)
    m.eval()
    out = m(input_ids=ids, labels=ids)
    assert rel(out["logits"], gold["logits0"]) <= 2e-2
    m.train()
    loss = m(input_ids=ids, labels=ids)["loss"]
    assert abs(loss.item() - gold["loss0"].item()) <= 5e-3 * gold["loss0"].item()
    loss.backward()
    grads = dict(m.named_parameters())
    worst = 0.0
    for k, g in gold["grads0"].items():
        if k.endswith("attention.self.key.bias"):
            # softmax is invariant to a per-query shift of the scores, so this gradient is analytically 0 (HF holds ~1e-10 of
            # rounding noise): check that ours is noise too, relative to the query-bias gradient of the same layer
            qn = gold["grads0"][k.replace("key.bias", "query.bias")].norm().item()
            assert grads[k].grad.float().norm().item() <= 2e-2 * qn, k
            continue
        e = rel(grads[k].grad, g)
        worst = max(worst, e)
        assert e <= 5e-2, f"{k}: rel err {e:.3e}"
    print("worst grad rel err", worst)


def test_three_adam_steps_vs_hf_golden(gold, dev):
    m = build(gold, dev)
    opt = B200Adam(m.parameters(), lr=4e-4, betas=(0.9, 0.98), weight_decay=0.0)
    losses = []
    for b in gold["batches"]:
        loss = m(input_ids=b.to(dev), labels=b.to(dev))["loss"]
        loss.backward()
        opt.step()
        m.zero_grad()
        losses.append(loss.item())
    ref = gold["losses_3steps"]
    for a, b in zip(losses, ref.tolist()):
        assert abs(a - b) <= 1e-2 * b, (losses, ref.tolist())
    sd = m.state_dict()
    for k in ("roberta.encoder.layer.1.intermediate.dense.weight", "roberta.embeddings.word_embeddings.weight", "lm_head.bias"):
        # Adam's first steps move every weight by ~lr regardless of gradient scale: compare the UPDATE, loosely
        upd, ref_upd = sd[k].cpu() - gold["state_dict"][k], gold["state_dict_after3"][k] - gold["state_dict"][k]
        assert rel(upd, ref_upd) <= 0.25, (k, rel(upd, ref_upd))
    # the vocabulary padding behind the embedding matrix / decoder bias must stay exactly zero
    f = m.flat
    assert torch.all(f.view_alloc(f.master, "roberta.embeddings.word_embeddings.weight")[m.V:] == 0)
    assert torch.all(f.view_alloc(f.master, "lm_head.bias")[m.V:] == 0)


def test_loss_and_grads_vs_oracle_roberta_large_head_shape(dev):
    """hidden 256 / 4 heads (head_dim 64, roberta-large's), odd vocabulary like 50265, S = 512."""
    cfg = dict(vocab_size=1001, hidden_size=256, num_hidden_layers=2, num_attention_heads=4, intermediate_size=1024,
               max_position_embeddings=514, type_vocab_size=1, pad_token_id=1, layer_norm_eps=1e-5, hidden_dropout_prob=0.0,
               attention_probs_dropout_prob=0.0, hidden_act="gelu", initializer_range=0.02)
    m = B200RobertaForMaskedLM(SimpleNamespace(**cfg))
    m.reset_parameters(torch.Generator().manual_seed(0))
    P = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.to(dev).train()
    ids = torch.randint(0, 1001, (2, 512), generator=torch.Generator().manual_seed(1))
    loss = m(input_ids=ids.to(dev), labels=ids.to(dev))["loss"]
    loss.backward()
    ref_loss, ref_grads, _ = R.roberta_loss_and_grads(P, ids, ids, cfg)
    assert abs(loss.item() - ref_loss.item()) <= 5e-3 * ref_loss.item()
    grads = dict(m.named_parameters())
    for k in ("roberta.encoder.layer.0.attention.self.query.weight", "roberta.encoder.layer.1.output.dense.weight",
              "roberta.embeddings.word_embeddings.weight", "roberta.embeddings.position_embeddings.weight", "lm_head.dense.weight"):
        assert rel(grads[k].grad, ref_grads[k]) <= 5e-2, (k, rel(grads[k].grad, ref_grads[k]))


def test_dropout_kernel_statistics_and_backward_mask(dev):
    n = 1 << 22
    x = torch.ones(n, device=dev, dtype=torch.bfloat16)
    y = K.dropout(x, 0.1, seed=1234)
    keep = (y != 0).float().mean().item()
    assert abs(keep - 0.9) < 2e-3, keep
    assert abs(y.float().mean().item() - 1.0) < 3e-3  # E[out] = in
    assert torch.equal(y, K.dropout(x, 0.1, seed=1234))  # pure function of (seed, index): backward recomputes it
    assert not torch.equal(y, K.dropout(x, 0.1, seed=1235))
    r = torch.randn(n, device=dev).to(torch.bfloat16)
    z = K.dropout(x, 0.1, seed=1234, residual=r)
    assert torch.allclose(z.float(), y.float() + r.float(), atol=2e-2)
    # neighbouring elements are uncorrelated
    k = (y != 0).float()
    c = ((k[:-1] - 0.9) * (k[1:] - 0.9)).mean().item() / (0.9 * 0.1)
    assert abs(c) < 5e-3, c


def test_training_with_hidden_dropout_runs_and_learns(gold, dev):
    m = build(gold, dev, hidden_dropout_prob=0.1)
    opt = B200Adam(m.parameters(), lr=1e-3, betas=(0.9, 0.98))
    b = gold["batches"][0].to(dev)
    first = last = None
    for _ in range(30):
        loss = m(input_ids=b, labels=b)["loss"]
        loss.backward()
        opt.step()
        m.zero_grad()
        first = loss.item() if first is None else first
        last = loss.item()
    assert torch.isfinite(torch.tensor(last)) and last < 0.7 * first, (first, last)


def test_training_with_attention_dropout_runs_and_learns(gold, dev):
    m = build(gold, dev, hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1)
    opt = B200Adam(m.parameters(), lr=1e-3, betas=(0.9, 0.98))
    b = gold["batches"][0].to(dev)
    first = last = None
    for _ in range(30):
        loss = m(input_ids=b, labels=b)["loss"]
        loss.backward()
        opt.step()
        m.zero_grad()
        first = loss.item() if first is None else first
        last = loss.item()
    assert torch.isfinite(torch.tensor(last)) and last < 0.7 * first, (first, last)
    # eval mode applies no dropout: deterministic
    m.eval()
    a = m(input_ids=b, labels=b)["loss"].item()
    assert a == m(input_ids=b, labels=b)["loss"].item()


@pytest.mark.parametrize("p_attn", [0.1, 0.0])
def test_full_depth_training_steps_with_dropout(dev, p_attn):
    """roberta-large depth and width (24 layers, hidden 1024, 16 heads, S 512) at micro-batch 8: every kernel of the step at
    its real per-layer shapes, 24 times in a row, forward + backward + optimizer, with and without attention dropout."""
    from multimodal_llm_pretraining_b200.models.configs import as_namespace, roberta_large_config_dict

    cfg = dict(roberta_large_config_dict(), attention_probs_dropout_prob=p_attn)
    torch.manual_seed(0)
    m = B200RobertaForMaskedLM(as_namespace(cfg)).to(dev).train()
    opt = B200Adam(m.parameters(), lr=1e-4, betas=(0.9, 0.98))
    ids = torch.randint(0, cfg["vocab_size"], (8, 512), generator=torch.Generator().manual_seed(1)).to(dev)
    losses = []
    for _ in range(3):
        loss = m(input_ids=ids, labels=ids)["loss"]
        loss.backward()
        opt.step()
        m.zero_grad()
        torch.cuda.synchronize()
        losses.append(loss.item())
    assert all(torch.isfinite(torch.tensor(losses))), losses
    assert abs(losses[0] - 10.9) < 0.5 and losses[-1] < losses[0], losses  # ln(50265) = 10.82 at random init; it learns


def test_activation_checkpointing_reproduces_gradients_with_dropout(gold, dev):
    """Recomputing a layer in backward must regenerate the same dropout masks (they are functions of (step seed, site,
    element)): gradients with and without per-layer checkpointing agree to the fp32-accumulation noise of the wgrad
    reductions (split-K partial sums meet through red.add in arbitrary order)."""
    ids = gold["batches"][0]
    grads = []
    for ckpt in (False, True):
        m = build(gold, dev, hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1)
        if ckpt:
            m.gradient_checkpointing_enable()
            assert m.is_gradient_checkpointing
        loss = m(input_ids=ids.to(dev), labels=ids.to(dev))["loss"]
        loss.backward()
        grads.append((loss.item(), m.flat.grad.clone()))
    assert grads[0][0] == grads[1][0]
    assert rel(grads[1][1], grads[0][1]) <= 1e-5


def test_gemm_fused_dropout_equals_standalone_dropout_kernel(dev):
    """dropout fused into the bias + residual GEMM epilogue must draw the mask of K.dropout(seed) (its backward is that kernel applied to
    the gradient): same kept set; kept values differ only by the bf16 rounding the unfused path adds between the GEMM and the dropout."""
    for M, N, Kd in [(512, 1024, 256), (200, 136, 64), (4096, 1024, 1024)]:
        g = torch.Generator(device="cpu").manual_seed(M)
        A = torch.randn(M, Kd, generator=g).to(dev).to(torch.bfloat16)
        W = (torch.randn(N, Kd, generator=g) * 0.05).to(dev).to(torch.bfloat16)
        b = torch.randn(N, generator=g).to(dev)
        res = torch.randn(M, N, generator=g).to(dev).to(torch.bfloat16)
        z = K.gemm(A, W, bias=b)
        want = K.dropout(z, 0.1, seed=77, residual=res)
        got = K.gemm(A, W, bias=b, residual=res, dropout_p=0.1, dropout_seed=77)
        dropped_ref = (want == res)
        dropped = (got == res)
        assert (dropped_ref != dropped).float().mean().item() < 1e-4  # coincidences where z rounds to 0 only
        frac = dropped.float().mean().item()
        assert abs(frac - 0.1) < 0.01, frac
        assert rel(got, want) <= 5e-3, rel(got, want)
        # backward of the fused op = the stand-alone kernel on the gradient with the same seed: zero exactly where the forward dropped
        # (the kept set read off a run with a zero residual, where "output == 0" can only mean "dropped")
        kept0 = K.gemm(A, W, bias=b, residual=torch.zeros_like(res), dropout_p=0.1, dropout_seed=77) != 0
        gr = torch.ones(M, N, device=dev, dtype=torch.bfloat16)
        gmask = K.dropout(gr, 0.1, seed=77)
        assert ((gmask != 0) != kept0).float().mean().item() < 1e-5
