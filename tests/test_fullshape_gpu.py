"""Full-shape parity on the B200 against the live fp32 HF modules (eager / sdpa, same weights, same token batch, TF32 off):
one real layer of every BASELINE.json model at its real width, head count, sequence length and micro-batch, plus the LM
head + cross entropy at M = 32 768 x V = 50 304 (the ragged 256-wide N tile: 50 304 = 196 x 256 + 128).

The toy-shape model tests (test_model_gpu.py) prove the wiring; these prove the kernels at the shapes the bench runs:
  Pythia-1b   layer  h 2048, 8 heads x 256,  rot 64, S 2048, B 16  -> T = 32 768 rows through every GEMM, hd-256 attention
  Pythia-2.8b layer  h 2560, 32 heads x 80,  rot 20, S 2048, B 4   -> padded head_dim-80 attention path
  Pythia-410m layer  h 1024, 16 heads x 64,  rot 16, S 2048, B 8   -> two-CTA hd-64 attention
  RoBERTa-large layer h 1024, 16 heads x 64, S 512 (B 16) and S 128 (B 64), post-LN, bidirectional, tied padded decoder
Tolerance (north_star): bf16 operands vs an fp32 reference => rel <= 2e-2 of the tensor's norm, for the loss 2e-3.
HF: models/gpt_neox/modeling_gpt_neox.py:258-289 (layer), :327-334 (head); models/roberta/modeling_roberta.py;
reference builder: src/models/pythia.py:15-22, src/models/roberta.py:15-18."""
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from multimodal_llm_pretraining_b200 import kernels as K  # noqa: E402
from multimodal_llm_pretraining_b200.modeling_gpt_neox import B200GPTNeoXForCausalLM  # noqa: E402
from multimodal_llm_pretraining_b200.modeling_roberta import B200RobertaForMaskedLM  # noqa: E402
from multimodal_llm_pretraining_b200.models.configs import as_namespace, pythia_config_dict, roberta_large_config_dict  # noqa: E402

TOL = 2e-2


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.fixture(autouse=True)
def _fp32_reference_math():
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    torch.cuda.empty_cache()


def _perturb_biases(hf, seed):
    """HF initialises biases to 0 and LayerNorm to (1, 0): give them values so that their forward use is exercised."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in hf.named_parameters():
            if n.endswith(".bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
            elif "layernorm.weight" in n.lower() or "layer_norm.weight" in n.lower():
                p.copy_(1.0 + torch.randn(p.shape, generator=g) * 0.05)


def _compare_grads(mine, hf, skip=()):
    theirs = dict(hf.named_parameters())
    worst = (0.0, None)
    checked = 0
    for n, p in mine.named_parameters():
        if any(s in n for s in skip):
            continue
        ref = theirs[n].grad
        assert ref is not None, n
        e = rel(p.grad, ref)
        worst = max(worst, (e, n))
        assert e <= TOL, f"{n}: grad rel err {e:.3e} > {TOL}"
        checked += 1
    assert checked > 0
    return worst


@pytest.mark.parametrize("name,B", [("pythia-1b", 16), ("pythia-2.8b", 4), ("pythia-410m", 8)])
def test_neox_layer_and_lm_head_full_shape_vs_hf_fp32(dev, name, B):
    tr = pytest.importorskip("transformers")
    cfg = dict(pythia_config_dict(name), num_hidden_layers=1)
    torch.manual_seed(0)
    hf = tr.GPTNeoXForCausalLM(tr.GPTNeoXConfig(**cfg, attn_implementation="sdpa")).float()
    _perturb_biases(hf, 1)
    sd = {k: v.detach().clone() for k, v in hf.state_dict().items() if "inv_freq" not in k}
    hf = hf.to(dev).train()
    mine = B200GPTNeoXForCausalLM(as_namespace(cfg))
    mine.load_hf_state_dict(sd)
    mine = mine.to(dev).train()
    ids = torch.randint(0, cfg["vocab_size"], (B, 2049), generator=torch.Generator().manual_seed(2)).to(dev)
    loss = mine(input_ids=ids, labels=ids).loss
    loss.backward()
    ref = hf(input_ids=ids, labels=ids).loss
    ref.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - ref.item()) <= 2e-3 * ref.item(), (loss.item(), ref.item())
    worst = _compare_grads(mine, hf)
    print(f"{name} 1 layer + head, B {B} x 2049: loss {loss.item():.5f} vs HF fp32 {ref.item():.5f}; worst grad rel err {worst[0]:.3e} ({worst[1]})")


@pytest.mark.parametrize("S,B", [(512, 16), (128, 64)])
def test_roberta_large_layer_full_shape_vs_hf_fp32(dev, S, B):
    tr = pytest.importorskip("transformers")
    cfg = dict(roberta_large_config_dict(), num_hidden_layers=1, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    torch.manual_seed(0)
    hf = tr.RobertaForMaskedLM(tr.RobertaConfig(**cfg, attn_implementation="eager")).float()
    _perturb_biases(hf, 3)
    sd = {k: v.detach().clone() for k, v in hf.state_dict().items() if "position_ids" not in k and "token_type_ids" not in k}
    hf = hf.to(dev).train()
    mine = B200RobertaForMaskedLM(as_namespace(cfg))
    mine.load_hf_state_dict(sd)
    mine = mine.to(dev).train()
    ids = torch.randint(0, cfg["vocab_size"], (B, S), generator=torch.Generator().manual_seed(4))
    ids[:, -3:] = 1  # trailing pad tokens: position ids and the padding_idx gradient rule are on the path
    ids = ids.to(dev)
    loss = mine(input_ids=ids, labels=ids)["loss"]
    loss.backward()
    ref = hf(input_ids=ids, labels=ids).loss
    ref.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - ref.item()) <= 2e-3 * ref.item(), (loss.item(), ref.item())
    # the key bias gradient is analytically zero (softmax is invariant to a per-query shift of the scores): noise on both sides
    worst = _compare_grads(mine, hf, skip=("attention.self.key.bias",))
    print(f"roberta-large 1 layer + head, {B} x {S}: loss {loss.item():.5f} vs HF fp32 {ref.item():.5f}; worst grad rel err {worst[0]:.3e} ({worst[1]})")


def test_lm_head_gemm_and_cross_entropy_at_32768_x_50304(dev):
    """The LM-head GEMM pair and the one-pass cross entropy alone, at the bench's M = 32 768, K = 2048, V = 50 304, against
    fp32 torch computed from the SAME bf16 operands (so only the kernels' own rounding is in the difference)."""
    T, h, V = 32768, 2048, 50304
    g = torch.Generator(device=dev).manual_seed(0)
    x = (torch.randn(T, h, device=dev, generator=g) * 1.0).to(torch.bfloat16)
    W = (torch.randn(V, h, device=dev, generator=g) * 0.02).to(torch.bfloat16)
    labels = torch.randint(0, V, (T,), device=dev, generator=g)
    labels[::97] = -100
    logits = K.gemm(x, W)
    # reference in row blocks (fp32 logits of the whole problem would be 6.6 GB; fine on 180 GB, but there is no need)
    n_valid = int((labels != -100).sum())
    ref_loss = torch.zeros((), device=dev, dtype=torch.float64)
    dW_ref = torch.zeros(V, h, device=dev, dtype=torch.float32)
    dx_ref = torch.empty(T, h, device=dev, dtype=torch.float32)
    worst_logits = 0.0
    Wf = W.float()
    for r0 in range(0, T, 4096):
        xs = x[r0:r0 + 4096].float()
        lg = xs @ Wf.t()
        worst_logits = max(worst_logits, rel(logits[r0:r0 + 4096], lg))
        lb = labels[r0:r0 + 4096]
        ref_loss += torch.nn.functional.cross_entropy(lg, lb, ignore_index=-100, reduction="sum").double()
        p = torch.softmax(lg, dim=-1)
        valid = lb != -100
        p[valid, lb[valid]] -= 1.0
        p[~valid] = 0.0
        p /= n_valid
        dW_ref += p.t() @ xs
        dx_ref[r0:r0 + 4096] = p @ Wf
    ref_loss = (ref_loss / n_valid).item()
    assert worst_logits <= 1e-2, worst_logits
    loss, nv = K.cross_entropy_(logits, labels, V=V, write_grad=True)  # logits now hold dlogits (bf16)
    assert int(nv.item()) == n_valid
    assert abs(loss.item() - ref_loss) <= 1e-3 * ref_loss, (loss.item(), ref_loss)
    dW = torch.zeros(V, h, device=dev, dtype=torch.float32)
    K.gemm(logits, x, a_mn=True, b_mn=True, out=dW, accumulate=True)
    dx = K.gemm(logits, W, b_mn=True)
    torch.cuda.synchronize()
    e_w, e_x = rel(dW, dW_ref), rel(dx, dx_ref)
    print(f"LM head 32768 x 50304: logits rel {worst_logits:.2e}, loss {loss.item():.5f} vs {ref_loss:.5f}, dW rel {e_w:.2e}, dx rel {e_x:.2e}")
    assert e_w <= TOL and e_x <= TOL, (e_w, e_x)
