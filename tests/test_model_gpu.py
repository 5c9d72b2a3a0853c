"""Module-level parity on the B200: B200GPTNeoXForCausalLM + B200Adam (C-ABI kernels) against
 (a) the golden vectors of the real HF module (tests/golden/neox_tiny.pt),
 (b) the fp32 CPU oracle on other shapes (head_dim 128 / 256),
 (c) the live HF module trained side by side for 200 steps (loss curve within 1 %, north_star).
Tolerances are vs fp32 references; bf16 operands => rel <= 2e-2 per tensor (north_star), whole-model parameter gradients included.
The one exception is stated where it applies: the key third of query_key_value.bias has an analytically ZERO gradient (softmax is
invariant to a per-query shift of the scores), so HF holds rounding noise there and so do we; it is compared on the query and value
thirds only. After three Adam steps the parameter UPDATE is compared: Adam's first steps move every element by ~lr * sign(g), so
elements whose gradient is smaller than the bf16 rounding of the backward pass take opposite +-lr steps; the bound per tensor is
stated in the test and the measured values are printed."""
import sys
from pathlib import Path
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from multimodal_llm_pretraining_b200.modeling_gpt_neox import B200GPTNeoXForCausalLM  # noqa: E402
from multimodal_llm_pretraining_b200.optim import B200Adam, get_scheduler  # noqa: E402
from oracle import neox_oracle as O  # noqa: E402

GOLD = ROOT / "tests" / "golden" / "neox_tiny.pt"


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def build(cfg: dict, dev, sd=None, seed=0):
    m = B200GPTNeoXForCausalLM(SimpleNamespace(**cfg))
    if sd is not None:
        m.load_hf_state_dict(sd)
    else:
        m.reset_parameters(torch.Generator().manual_seed(seed))
        with torch.no_grad():
            g = torch.Generator().manual_seed(seed + 1)
            for n, p in m.named_parameters():
                if n.endswith(".bias"):
                    p.normal_(0, 0.02, generator=g)
    return m.to(dev).train()


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, map_location="cpu", weights_only=False)


def test_state_dict_keys_match_hf(gold, dev):
    m = build(gold["cfg"], dev, gold["state_dict"])
    assert sorted(m.state_dict().keys()) == sorted(gold["state_dict"].keys())
    for k, v in m.state_dict().items():
        assert v.shape == gold["state_dict"][k].shape and v.dtype == torch.float32


def test_loss_and_grads_vs_hf_golden(gold, dev):
    m = build(gold["cfg"], dev, gold["state_dict"])
    ids = gold["batches"][0].to(dev)
    out = m(input_ids=ids, labels=ids)
    assert out.get("loss") is out["loss"] is out.loss
    assert abs(out.loss.item() - gold["loss0"]) <= 2e-3 * gold["loss0"], (out.loss.item(), gold["loss0"])
    out.loss.backward()
    grads = {n: p.grad for n, p in m.named_parameters()}
    worst = (0.0, None)
    for k, g in gold["grads"].items():
        a, b = grads[k], g
        if k.endswith("query_key_value.bias"):  # [nh, 3, hd]: drop the key third (analytically zero gradient, see module docstring)
            nh = gold["cfg"]["num_attention_heads"]
            a, b = a.view(nh, 3, -1)[:, [0, 2]], g.view(nh, 3, -1)[:, [0, 2]]
        worst = max(worst, (rel(a, b), k))
        assert rel(a, b) <= 2e-2, (k, rel(a, b))
    print("whole-model gradients vs HF fp32 golden: worst rel err %.3e (%s), tolerance 2e-2" % worst)
    for k, n in gold["grad_norms"].items():
        got = grads[k].norm().item()
        assert abs(got - n) <= 5e-2 * n + 1e-6, (k, got, n)


def test_three_adam_steps_vs_hf_golden(gold, dev):
    m = build(gold["cfg"], dev, gold["state_dict"])
    decay = [p for n, p in m.named_parameters() if p.dim() >= 2]
    no_decay = [p for n, p in m.named_parameters() if p.dim() < 2]
    opt = B200Adam([{"params": decay, "weight_decay": 0.0}, {"params": no_decay, "weight_decay": 0.0}], lr=6e-4, betas=(0.9, 0.95), eps=1e-8)
    losses, norms = [], []
    for ids in gold["batches"]:
        ids = ids.to(dev)
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
    errs = []
    for k, v in gold["params_after3"].items():
        # 3 Adam steps move each weight by ~3*lr; compare the UPDATE, not the weight
        upd_ref = v - gold["state_dict"][k]
        upd_got = sd[k].cpu() - gold["state_dict"][k]
        if k.endswith("query_key_value.bias"):
            nh = gold["cfg"]["num_attention_heads"]
            upd_ref, upd_got = upd_ref.view(nh, 3, -1)[:, [0, 2]], upd_got.view(nh, 3, -1)[:, [0, 2]]
        errs.append((rel(upd_got, upd_ref), k))
    errs.sort(reverse=True)
    print("3-step Adam update vs HF fp32 golden, worst tensors:", [(k, round(e, 4)) for e, k in errs[:5]])
    for e, k in errs:
        # sign flips of near-zero gradients under Adam's ~lr*sign(g) first steps; a wrong update rule or a stale moment gives O(1)
        assert e <= 0.1, (k, e)


@pytest.mark.parametrize("nh,h", [(2, 256), (1, 256), (4, 320)])  # head_dim 128, 256, 80 (Pythia-1.4b / 1b / 2.8b head shapes)
def test_loss_and_grads_vs_oracle_other_head_dims(dev, nh, h):
    cfg = dict(vocab_size=512, hidden_size=h, num_hidden_layers=2, num_attention_heads=nh, intermediate_size=4 * h,
               rotary_pct=0.25, rotary_emb_base=10000, layer_norm_eps=1e-5)
    m = build(cfg, dev, seed=nh)
    ids = torch.randint(0, 512, (2, 193), generator=torch.Generator().manual_seed(5))
    P = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    ref_loss, ref_grads = O.neox_loss_and_grads(P, ids, ids, cfg)
    loss = m(input_ids=ids.to(dev), labels=ids.to(dev)).loss
    loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 2e-3 * ref_loss.item()
    worst = (0.0, None)
    for n, p in m.named_parameters():
        a, b = p.grad, ref_grads[n]
        if n.endswith("query_key_value.bias"):
            a, b = a.view(nh, 3, -1)[:, [0, 2]], b.view(nh, 3, -1)[:, [0, 2]]
        worst = max(worst, (rel(a, b), n))
        assert rel(a, b) <= 2e-2, (n, rel(a, b))
    print("head_dim %d gradients vs the fp32 oracle: worst rel err %.3e (%s), tolerance 2e-2" % ((h // nh,) + worst))


def test_grad_accumulation_checkpointing_and_torch_optimizer_interop(gold, dev):
    m = build(gold["cfg"], dev, gold["state_dict"])
    ids0, ids1 = gold["batches"][0].to(dev), gold["batches"][1].to(dev)
    # accumulate two micro-batches with loss/2 (HF Trainer convention, HF:trainer.py:1925-1927)
    for ids in (ids0, ids1):
        (m(input_ids=ids, labels=ids).loss / 2).backward()
    acc = m.flat.grad.clone()
    m.zero_grad()
    assert float(m.flat.grad.abs().max()) == 0.0
    m(input_ids=ids0, labels=ids0).loss.backward()
    g0 = m.flat.grad.clone()
    m.zero_grad()
    m(input_ids=ids1, labels=ids1).loss.backward()
    g1 = m.flat.grad.clone()
    assert rel(acc, (g0 + g1) / 2) <= 2e-3
    # activation checkpointing recomputes the same kernels -> same gradients
    m.zero_grad()
    m.gradient_checkpointing_enable()
    m(input_ids=ids1, labels=ids1).loss.backward()
    assert rel(m.flat.grad, g1) <= 1e-5
    m.gradient_checkpointing_disable()
    # a stock torch optimizer (the reference's own class) works on the same parameters; set_to_none is survived
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    opt.step()
    opt.zero_grad(set_to_none=True)
    before = m.flat.shadow.clone()
    loss = m(input_ids=ids0, labels=ids0).loss  # re-attaches grads, re-syncs the bf16 shadow
    assert not torch.equal(before, m.flat.shadow)
    loss.backward()
    assert all(p.grad is not None for p in m.parameters())
    # eval / inference paths
    m.eval()
    with torch.no_grad():
        out = m(input_ids=ids0)
    assert out.logits.shape == (ids0.shape[0], ids0.shape[1], gold["cfg"]["vocab_size"])
    ref_logits = O.neox_logits({k: v.cpu() for k, v in m.state_dict().items()}, ids0.cpu(), gold["cfg"])
    assert rel(out.logits, ref_logits) <= 2e-2


def test_loss_curve_200_steps_vs_hf(dev):
    """north_star: loss curves within 1 % over 200 steps, identical init, identical synthetic token batches."""
    tr = pytest.importorskip("transformers")
    cfg = dict(vocab_size=1024, hidden_size=256, num_hidden_layers=4, num_attention_heads=4, intermediate_size=1024,
               max_position_embeddings=256, rotary_pct=0.25, rotary_emb_base=10000, layer_norm_eps=1e-5,
               use_parallel_residual=True, hidden_act="gelu", attention_bias=True, hidden_dropout=0.0,
               attention_dropout=0.0, tie_word_embeddings=False, initializer_range=0.02)
    torch.manual_seed(0)
    hf = tr.GPTNeoXForCausalLM(tr.GPTNeoXConfig(**cfg, attn_implementation="sdpa")).float().to(dev).train()
    sd = {k: v.detach().cpu().clone() for k, v in hf.state_dict().items() if "inv_freq" not in k}
    mine = build(cfg, dev, sd)
    g = torch.Generator().manual_seed(3)
    data = [torch.randint(0, 1024, (8, 129), generator=g).to(dev) for _ in range(16)]  # small fixed corpus: loss must fall
    steps, warm = 200, 20
    kw = dict(lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.0)
    o_hf = torch.optim.Adam(hf.parameters(), **kw)
    o_me = B200Adam(mine.parameters(), **kw)
    s_hf = get_scheduler("cosine_with_min_lr", o_hf, warm, steps, {"min_lr_rate": 0.1})
    s_me = get_scheduler("cosine_with_min_lr", o_me, warm, steps, {"min_lr_rate": 0.1})
    l_hf, l_me = [], []
    for t in range(steps):
        ids = data[t % len(data)]
        a = hf(input_ids=ids, labels=ids).loss
        a.backward()
        torch.nn.utils.clip_grad_norm_(hf.parameters(), 1.0)
        o_hf.step(), s_hf.step(), hf.zero_grad()
        b = mine(input_ids=ids, labels=ids).loss
        b.backward()
        mine.clip_grad_norm_(1.0)
        o_me.step(), s_me.step(), mine.zero_grad()
        l_hf.append(a.item()), l_me.append(b.item())
    assert l_hf[-1] < 0.9 * l_hf[0], f"reference did not learn: {l_hf[0]} -> {l_hf[-1]}"
    worst = max(abs(x - y) / y for x, y in zip(l_me, l_hf))
    print(f"loss curve: hf {l_hf[0]:.4f}->{l_hf[-1]:.4f}  b200 {l_me[0]:.4f}->{l_me[-1]:.4f}  worst rel dev {worst:.4%}")
    assert worst <= 1e-2, f"worst relative deviation {worst:.3%} over 200 steps"
