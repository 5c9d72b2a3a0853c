"""CPU tests of the HOST side of the hot path with the REAL modules: B200GPTNeoXForCausalLM / B200RobertaForMaskedLM's hand-scheduled
forward and backward, the flat store's routing, B200Adam's chunk tables and the TrainEngine strategies run here on CPU tensors with the
C-ABI wrappers replaced by torch statements of their contracts (tests/cpu_kernels.py, test infrastructure). What is checked is
everything ABOVE the kernels: which tensor feeds which GEMM in which layout, where every gradient is accumulated, what the
optimizer touches, what the engine exchanges — against the HF golden fixtures and against each other across strategies (gloo,
world_size 2). The kernels themselves are checked on the GPU (`-m gpu`); the product's own CPU guards stay in place (the test
subclasses override `_require_cuda`, nothing else)."""
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
