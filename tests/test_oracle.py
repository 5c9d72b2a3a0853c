"""CPU tests (run with -m "not gpu"): the oracle restatement is pinned against golden vectors produced by the real
reference call path (transformers.GPTNeoXForCausalLM + torch.optim.Adam; tests/golden/make_golden.py)."""
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import neox_oracle as O  # noqa: E402

GOLD = ROOT / "tests" / "golden" / "neox_tiny.pt"


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, map_location="cpu", weights_only=False)


def test_oracle_loss_and_logits_match_hf_golden(gold):
    P = {k: v.clone() for k, v in gold["state_dict"].items()}
    ids = gold["batches"][0]
    logits = O.neox_logits(P, ids, gold["cfg"])
    assert torch.allclose(logits[:, :4, :8], gold["logits0_slice"], atol=2e-5, rtol=1e-4)
    loss = O.causal_lm_loss(logits, ids)
    assert abs(loss.item() - gold["loss0"]) < 2e-5


def test_oracle_grads_match_hf_golden(gold):
    P = {k: v.clone() for k, v in gold["state_dict"].items()}
    ids = gold["batches"][0]
    _, grads = O.neox_loss_and_grads(P, ids, ids, gold["cfg"])
    for k, n in gold["grad_norms"].items():
        assert abs(grads[k].norm().item() - n) <= 1e-4 * max(n, 1e-3), k
    for k, g in gold["grads"].items():
        assert torch.allclose(grads[k], g, atol=1e-6, rtol=1e-3), k


def test_oracle_adam_steps_match_torch_golden(gold):
    P = {k: v.clone() for k, v in gold["state_dict"].items()}
    losses = O.train_steps(P, gold["batches"], gold["cfg"], lr=6e-4, betas=(0.9, 0.95), eps=1e-8, max_grad_norm=1.0,
                           warmup=0, total_steps=1)
    for a, b in zip(losses, gold["losses"]):
        assert abs(a - b) < 5e-5, (losses, gold["losses"])
    for k, v in gold["params_after3"].items():
        assert torch.allclose(P[k], v, atol=2e-6, rtol=1e-4), k
    for k, n in gold["param_norms_after3"].items():
        assert abs(P[k].norm().item() - n) <= 1e-5 * max(n, 1e-3), k


def test_dead_last_token_identity(gold):
    """model(ids)[loss] with HF's internal shift == model(ids[:, :-1]) against ids[:, 1:]."""
    P = {k: v.double() for k, v in gold["state_dict"].items()}
    ids = gold["batches"][1]
    full = O.neox_loss(P, ids, ids, gold["cfg"])
    logits = O.neox_logits(P, ids[:, :-1], gold["cfg"])
    short = torch.nn.functional.cross_entropy(logits.reshape(-1, logits.shape[-1]), ids[:, 1:].reshape(-1))
    assert abs(full.item() - short.item()) < 1e-6  # HF upcasts the loss path to fp32


def test_oracle_matches_live_hf_when_available(gold):
    tr = pytest.importorskip("transformers")
    cfg = tr.GPTNeoXConfig(**gold["cfg"], attn_implementation="eager")
    m = tr.GPTNeoXForCausalLM(cfg).float()
    m.load_state_dict(gold["state_dict"], strict=False)
    ids = gold["batches"][2]
    ref = m(input_ids=ids, labels=ids).loss
    got = O.neox_loss({k: v.clone() for k, v in gold["state_dict"].items()}, ids, ids, gold["cfg"])
    assert abs(ref.item() - got.item()) < 2e-5


@pytest.mark.parametrize("hidden,heads", [(128, 2), (160, 2), (256, 2), (256, 1)], ids=["hd64", "hd80", "hd128", "hd256"])
def test_oracle_matches_live_hf_at_every_head_shape(gold, hidden, heads):
    """The GPU parity tests at head_dim 64 / 80 / 128 / 256 (the shapes of Pythia-160m..410m, 2.8b, 1.4b, 1b; rotary 16 / 20 /
    32 / 64 dims) check the kernels against this oracle: pin it to the real HF module at each of those head shapes — loss
    and every parameter gradient."""
    tr = pytest.importorskip("transformers")
    cfg = dict(gold["cfg"], hidden_size=hidden, num_attention_heads=heads, intermediate_size=4 * hidden, vocab_size=96)
    torch.manual_seed(hidden + heads)
    m = tr.GPTNeoXForCausalLM(tr.GPTNeoXConfig(**cfg, attn_implementation="eager")).float()
    with torch.no_grad():  # non-trivial biases / LayerNorm parameters (HF initialises them to 0 / 1)
        for n, p in m.named_parameters():
            if p.dim() == 1:
                p.add_(torch.randn_like(p) * 0.05)
    ids = torch.randint(0, 96, (2, 33))
    ref = m(input_ids=ids, labels=ids).loss
    ref.backward()
    P = {k: v.detach().clone() for k, v in m.state_dict().items() if "inv_freq" not in k and "masked_bias" not in k and not k.endswith("attention.bias")}
    loss, grads = O.neox_loss_and_grads(P, ids, ids, cfg)
    assert abs(loss.item() - ref.item()) < 2e-5
    for n, p in m.named_parameters():
        assert torch.allclose(grads[n], p.grad, atol=2e-6, rtol=2e-4), n


def test_schedule_matches_hf_formula():
    # Pythia: warmup 1430 of 143000 steps, min_lr_rate 0.1 (src/models/pythia.py:69-78)
    assert O.cosine_with_min_lr(0, 1430, 143000, 0.1) == 0.0
    assert abs(O.cosine_with_min_lr(715, 1430, 143000, 0.1) - 0.5) < 1e-12
    assert abs(O.cosine_with_min_lr(1430, 1430, 143000, 0.1) - 1.0) < 1e-12
    assert abs(O.cosine_with_min_lr(143000, 1430, 143000, 0.1) - 0.1) < 1e-12
    from multimodal_llm_pretraining_b200.optim import cosine_with_min_lr_lambda
    for s in (0, 10, 1429, 1430, 50000, 143000):
        assert abs(cosine_with_min_lr_lambda(s, num_warmup_steps=1430, num_training_steps=143000, min_lr_rate=0.1)
                   - O.cosine_with_min_lr(s, 1430, 143000, 0.1)) < 1e-12


# ---------------------------------------------------------------------------------------------- RoBERTa oracle
@pytest.fixture(scope="module")
def rgold():
    return torch.load(Path(__file__).resolve().parent / "golden" / "roberta_tiny.pt", weights_only=False)


def test_roberta_oracle_matches_hf_golden(rgold):
    from oracle import roberta_oracle as R

    ids = rgold["batches"][0]
    loss, grads, logits = R.roberta_loss_and_grads(rgold["state_dict"], ids, ids, rgold["cfg"])
    assert torch.allclose(logits, rgold["logits0"], atol=2e-5, rtol=1e-5)
    assert abs(loss.item() - rgold["loss0"].item()) <= 1e-5
    for k, g in rgold["grads0"].items():
        assert torch.allclose(grads[k], g, atol=1e-6, rtol=1e-4), k


def test_roberta_oracle_matches_live_hf_at_roberta_large_head_shape(rgold):
    """head_dim 64 (roberta-large: 16 heads x 64), non-trivial biases, pad tokens inside the batch."""
    tr = pytest.importorskip("transformers")
    from oracle import roberta_oracle as R

    cfg = dict(rgold["cfg"], hidden_size=128, num_attention_heads=2, intermediate_size=512)
    torch.manual_seed(7)
    m = tr.RobertaForMaskedLM(tr.RobertaConfig(**cfg, attn_implementation="eager")).float().eval()  # eval: dropout off
    with torch.no_grad():
        for n, p in m.named_parameters():
            if p.dim() == 1:
                p.add_(torch.randn_like(p) * 0.05)
    ids = torch.randint(3, cfg["vocab_size"], (2, 40))
    ids[0, 5] = ids[1, 0] = ids[1, 39] = cfg["pad_token_id"]  # positions skip pad tokens (create_position_ids_from_input_ids)
    ref = m(input_ids=ids, labels=ids)
    ref.loss.backward()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    loss, grads, logits = R.roberta_loss_and_grads(sd, ids, ids, cfg)
    assert abs(loss.item() - ref.loss.item()) < 2e-5
    assert torch.allclose(logits, ref.logits, atol=5e-5, rtol=1e-4)
    for n, p in m.named_parameters():
        if n in grads and p.grad is not None:
            assert torch.allclose(grads[n], p.grad, atol=2e-6, rtol=2e-4), n


def test_roberta_position_ids_rule(rgold):
    from oracle import roberta_oracle as R

    ids = torch.tensor([[0, 5, 1, 7, 1, 1, 9], [1, 1, 4, 4, 4, 1, 2]])
    assert R.position_ids(ids, 1).tolist() == [[2, 3, 1, 4, 1, 1, 5], [1, 1, 2, 3, 4, 1, 5]]
