"""fp16 + dynamic loss scaling on the B200: the reference's actual precision for every Pythia but 1b and
for RoBERTa (src/models/pythia.py:33-41, src/models/roberta.py:29-30; DeepSpeed fp16 block src/train.py:143-150).

 * libb200pt_fp16.so (the same sources, IEEE-half element type) against the fp32 PyTorch statements of the ops: the bf16
   kernel tests are re-run with fp16 operands (tolerances unchanged or tighter: fp16 carries 3 more mantissa bits);
 * the device-side loss-scale update against a Python restatement of torch.amp.GradScaler._amp_update_scale_ and of
   DeepSpeed's DynamicLossScaler.update_scale (hysteresis), over scripted overflow sequences, bit for bit;
 * the overflow path end to end: an inf gradient skips the step (parameters, moments, step count untouched), halves the
   scale; clean steps match the bf16-free arithmetic of torch.optim.Adam;
 * an fp16 model step against live fp32 HF (loss, gradients after unscaling) and a 410m-shaped layer.
"""
import math
import os
import subprocess
import sys
from pathlib import Path
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from multimodal_llm_pretraining_b200 import kernels as K  # noqa: E402
from multimodal_llm_pretraining_b200.engine import LossScaler, TrainEngine  # noqa: E402
from multimodal_llm_pretraining_b200.modeling_gpt_neox import B200GPTNeoXForCausalLM  # noqa: E402
from multimodal_llm_pretraining_b200.optim import B200Adam  # noqa: E402

FP16 = torch.float16


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


# ---------------------------------------------------------------------------------------------- the fp16 kernel library
@pytest.fixture
def half_ops(monkeypatch):
    """Re-point the bf16 kernel tests at fp16 operands (they cast their inputs through a module-level constant)."""
    import test_attention_gpu as TA
    import test_kernels_gpu as TK

    monkeypatch.setattr(TK, "BF16", FP16)
    monkeypatch.setattr(TA, "BF16", FP16)
    return TK, TA


def test_fp16_library_is_a_different_build(dev):
    from multimodal_llm_pretraining_b200 import _lib

    a, b = _lib.lib_for(dev, torch.bfloat16), _lib.lib_for(dev, FP16)
    assert a.b200_elem_dtype() == 0 and b.b200_elem_dtype() == 1 and a is not b
    x = torch.tensor([1.0 + 2 ** -9, 3.0e-6, 70000.0], device=dev).repeat(4)
    y16, yb = torch.empty(12, dtype=FP16, device=dev), torch.empty(12, dtype=torch.bfloat16, device=dev)
    K.cast_f32_to_bf16(x, y16)
    K.cast_f32_to_bf16(x, yb)
    assert torch.equal(y16, x.to(FP16)) and torch.equal(yb, x.to(torch.bfloat16))
    assert torch.isinf(y16[2]) and not torch.isinf(yb[2])  # the range difference that makes the loss scale necessary


def test_fp16_layernorm_gelu_rope_embedding(dev, half_ops):
    TK, _ = half_ops
    TK.test_layernorm_fwd_bwd(dev, 128, 2048, True)
    TK.test_layernorm_fwd_bwd(dev, 257, 1024, False)
    TK.test_gelu(dev)
    TK.test_rope(dev, 16, 64, 16)
    TK.test_rope(dev, 4, 80, 20)
    TK.test_embedding(dev)


def test_fp16_gemm(dev, half_ops):
    TK, _ = half_ops
    for a_mn, b_mn in [(False, False), (False, True), (True, True), (True, False)]:
        TK.test_gemm_layouts(dev, a_mn, b_mn, 384, 768, 1024)
        TK.test_gemm_layouts(dev, a_mn, b_mn, 200, 136, 328)
    TK.test_gemm_layouts(dev, False, False, 4096, 2048, 2048)
    TK.test_gemm_epilogues(dev, 512, 1024, 512)
    TK.test_gemm_epilogues(dev, 300, 520, 192)
    TK.test_gemm_wgrad_shape(dev)


def test_fp16_cross_entropy_with_loss_scale(dev, half_ops):
    TK, _ = half_ops
    TK.test_cross_entropy(dev, 64, 50304, 50304)
    TK.test_cross_entropy(dev, 37, 50265, 50304)
    # the loss scale is folded into dlogits: without it (softmax - onehot) / n_valid underflows fp16 at real batch sizes
    T, V = 4096, 50304
    g = torch.Generator(device=dev).manual_seed(0)
    logits = (torch.randn(T, V, device=dev, generator=g) * 2).to(FP16)
    labels = torch.randint(0, V, (T,), device=dev, generator=g)
    ref_in = logits.float().requires_grad_(True)
    torch.nn.functional.cross_entropy(ref_in, labels).backward()
    scale = torch.full((1,), 65536.0, device=dev)
    scaled = logits.clone()
    K.cross_entropy_(scaled, labels, V=V, grad_scale=scale)
    e_scaled = rel(scaled.float() / 65536.0, ref_in.grad)
    plain = logits.clone()
    K.cross_entropy_(plain, labels, V=V)
    e_plain = rel(plain.float(), ref_in.grad)
    assert e_scaled <= 2e-3, e_scaled
    assert e_plain > 10 * e_scaled, (e_plain, e_scaled)  # subnormal / flushed gradients: this is what the scale prevents


@pytest.mark.parametrize("case", [(2, 256, 2, 64, True, True), (1, 512, 3, 64, False, True), (2, 384, 2, 128, True, True),
                                  (2, 256, 2, 256, True, True), (1, 2048, 2, 256, True, True), (2, 256, 4, 80, True, True)])
def test_fp16_attention(dev, half_ops, case):
    _, TA = half_ops
    TA.test_attention_fwd_bwd(dev, *case)


def test_fp16_adam_shadow(dev):
    n = 65536 * 3 + 64
    p = torch.randn(n, device=dev)
    g = torch.randn(n, device=dev)
    m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    sh = torch.zeros(n, dtype=FP16, device=dev)
    cs = torch.arange(0, n, 65536, dtype=torch.int64, device=dev)
    cl = torch.tensor([min(65536, n - s) for s in range(0, n, 65536)], dtype=torch.int32, device=dev)
    cg = torch.zeros(cs.numel(), dtype=torch.int32, device=dev)
    ref = p.clone().requires_grad_(True)
    ref.grad = g.clone()
    torch.optim.Adam([ref], lr=1e-2, betas=(0.9, 0.95)).step()
    groups = [dict(lr=1e-2, beta1=0.9, beta2=0.95, eps=1e-8, weight_decay=0.0, bias_corr1=0.1, bias_corr2=0.05, adamw_mode=False)]
    K.adam_step(p, g, m, v, sh, 0, cs, cl, cg, groups)
    assert rel(p, ref.detach()) <= 2e-6
    assert torch.equal(sh, p.to(FP16))


# ---------------------------------------------------------------------------------------------- norms, overflow flag, scaler
def test_sumsq_is_deterministic_and_chunked_variant_matches(dev):
    x = torch.randn(40_000_003, device=dev)
    outs = []
    for _ in range(4):
        o = torch.zeros((), device=dev)
        K.sumsq_(x, o)
        outs.append(o.item())
    assert len(set(outs)) == 1, outs  # bit-identical every time: DDP replicas compute identical clip coefficients
    ref = x.double().pow(2).sum().item()
    assert abs(outs[0] - ref) <= 2e-6 * ref
    # chunk-table variant over disjoint slices (what a ZeRO rank owns)
    starts = [0, 1_000_000, 7_777_776, 20_000_000]
    lens = [65536, 65536, 40000, 13]
    cs = torch.tensor(starts, dtype=torch.int64, device=dev)
    cl = torch.tensor(lens, dtype=torch.int32, device=dev)
    o = torch.full((), 5.0, device=dev)
    K.sumsq_chunks_(x, cs, cl, o)
    ref = 5.0 + sum(x[s:s + n].double().pow(2).sum().item() for s, n in zip(starts, lens))
    assert abs(o.item() - ref) <= 2e-6 * ref
    o2 = torch.full((), 5.0, device=dev)
    K.sumsq_chunks_(x, cs, cl, o2)
    assert o.item() == o2.item()


def test_clip_coef_unscales_and_flags_overflow(dev):
    ss = torch.tensor(4.0 * 65536.0 ** 2, device=dev)  # gradients of true norm 2 carrying a 2^16 loss scale
    scale = torch.full((1,), 65536.0, device=dev)
    flag = torch.full((1,), 7, dtype=torch.int32, device=dev)
    norm, coef = K.clip_coef(ss, 1.0, loss_scale=scale, found_inf=flag)
    assert abs(norm.item() - 2.0) < 1e-6 and flag.item() == 0
    assert abs(coef.item() - (1.0 / (2.0 + 1e-6)) / 65536.0) < 1e-12
    _, coef = K.clip_coef(ss, 0.0, loss_scale=scale, found_inf=flag)  # clipping off: pure unscale
    assert coef.item() == 1.0 / 65536.0
    for bad in (float("inf"), float("nan")):
        K.clip_coef(torch.tensor(bad, device=dev), 1.0, loss_scale=scale, found_inf=flag)
        assert flag.item() == 1
    # one inf element in a gradient buffer is enough
    gbuf = torch.randn(1_000_000, device=dev)
    gbuf[123_457] = float("inf")
    o = torch.zeros((), device=dev)
    K.sumsq_(gbuf, o)
    K.clip_coef(o, 1.0, loss_scale=scale, found_inf=flag)
    assert flag.item() == 1


def _scaler_oracle(kind, seq, init=2.0 ** 16):
    """torch.amp.GradScaler._amp_update_scale_ (kind 'torch': growth 2, backoff 0.5, interval given) and DeepSpeed's
    DynamicLossScaler.update_scale (kind 'deepspeed': delayed_shift = hysteresis 2, scale_window, min_scale 1), in Python."""
    scale, out = init, []
    if kind == "torch":
        interval, tracker = 4, 0
        for inf in seq:
            if inf:
                scale, tracker = scale * 0.5, 0
            else:
                tracker += 1
                if tracker == interval:
                    scale, tracker = scale * 2.0, 0
            out.append(scale)
    else:
        window, hyst, cur_hyst, last_overflow, it = 4, 2, 2, -1, 0
        for inf in seq:
            if inf:
                if hyst == 1 or cur_hyst == 1:
                    scale = max(scale / 2.0, 1.0)
                else:
                    cur_hyst -= 1
                last_overflow = it
            else:
                if (it - last_overflow) % window == 0:
                    cur_hyst = hyst
                    scale *= 2.0
            it += 1
            out.append(scale)
    return out


@pytest.mark.parametrize("kind", ["torch", "deepspeed"])
def test_loss_scale_update_matches_reference_semantics(dev, kind):
    seq = [0, 0, 0, 0, 1, 1, 1, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1]
    want = _scaler_oracle(kind, seq, init=2.0 ** 16 if kind == "torch" else 8.0)
    sc = LossScaler(dev, kind=kind, growth_interval=4, init_scale=2.0 ** 16 if kind == "torch" else 8.0)
    got = []
    for inf in seq:
        sc.found_inf.fill_(inf)
        sc.update()
        got.append(sc.scale.item())
    assert got == want, (got, want)
    if kind == "deepspeed":
        assert min(got) >= 1.0  # min_loss_scale


def _tiny(dev, dtype, seed=0):
    cfg = dict(vocab_size=512, hidden_size=256, num_hidden_layers=2, num_attention_heads=4, intermediate_size=1024,
               rotary_pct=0.25, rotary_emb_base=10000, layer_norm_eps=1e-5)
    m = B200GPTNeoXForCausalLM(SimpleNamespace(**cfg))
    m.reset_parameters(torch.Generator().manual_seed(seed))
    m.set_compute_dtype(dtype)
    return m.to(dev).train(), cfg


def test_overflow_step_is_skipped_and_scale_backs_off(dev):
    m, cfg = _tiny(dev, FP16)
    opt = B200Adam(m.parameters(), lr=1e-3, betas=(0.9, 0.95))
    eng = TrainEngine(m, opt, None, max_grad_norm=1.0, gradient_accumulation_steps=1, strategy="none")
    assert eng.loss_scaler is not None and m.loss_scale is eng.loss_scaler.scale and eng.loss_scaler.scale.item() == 65536.0
    ids = torch.randint(0, 512, (2, 129), generator=torch.Generator().manual_seed(1)).to(dev)
    eng.manual_training_step({"input_ids": ids, "labels": ids})
    assert eng.manual_optimization_step() is True
    before = m.flat.master.clone()
    mom = opt._m.clone()
    step_before = opt._step
    eng.manual_training_step({"input_ids": ids, "labels": ids})
    m.flat.grad[12345] = float("inf")  # an overflowed gradient element
    assert eng.manual_optimization_step() is False
    assert torch.equal(m.flat.master, before) and torch.equal(opt._m, mom) and opt._step == step_before
    assert eng.loss_scaler.scale.item() == 32768.0 and eng.loss_scaler.skipped_steps == 1
    assert float(m.flat.grad.abs().max()) == 0.0  # gradients are cleared either way
    eng.manual_training_step({"input_ids": ids, "labels": ids})
    assert eng.manual_optimization_step() is True and not torch.equal(m.flat.master, before)


def test_fp16_model_step_matches_hf_fp32_after_unscaling(dev):
    tr = pytest.importorskip("transformers")
    cfg = dict(vocab_size=1024, hidden_size=256, num_hidden_layers=3, num_attention_heads=4, intermediate_size=1024,
               max_position_embeddings=256, rotary_pct=0.25, rotary_emb_base=10000, layer_norm_eps=1e-5,
               use_parallel_residual=True, hidden_act="gelu", attention_bias=True, hidden_dropout=0.0,
               attention_dropout=0.0, tie_word_embeddings=False, initializer_range=0.02)
    torch.manual_seed(0)
    hf = tr.GPTNeoXForCausalLM(tr.GPTNeoXConfig(**cfg, attn_implementation="sdpa")).float().to(dev).train()
    sd = {k: v.detach().cpu().clone() for k, v in hf.state_dict().items() if "inv_freq" not in k}
    mine = B200GPTNeoXForCausalLM(SimpleNamespace(**cfg))
    mine.load_hf_state_dict(sd)
    mine.set_compute_dtype(FP16)
    mine = mine.to(dev).train()
    scaler = LossScaler(dev, kind="deepspeed")
    mine.loss_scale = scaler.scale
    ids = torch.randint(0, 1024, (8, 193), generator=torch.Generator().manual_seed(3)).to(dev)
    loss = mine(input_ids=ids, labels=ids).loss
    loss.backward()
    ref = hf(input_ids=ids, labels=ids).loss
    ref.backward()
    assert mine.flat.shadow.dtype == FP16
    assert abs(loss.item() - ref.item()) <= 2e-3 * ref.item()  # the reported loss is NOT scaled
    theirs = dict(hf.named_parameters())
    worst = 0.0
    for n, p in mine.named_parameters():
        e = rel(p.grad / 65536.0, theirs[n].grad)  # flat.grad holds gradients of 2^16 * loss
        worst = max(worst, e)
        assert e <= 2e-2, (n, e)
    print("fp16 vs HF fp32: worst grad rel err", worst)
    # Adam with the unscale folded into the clip coefficient == torch.optim.Adam on the true gradients
    o_hf = torch.optim.Adam(hf.parameters(), lr=1e-3, betas=(0.9, 0.95))
    torch.nn.utils.clip_grad_norm_(hf.parameters(), 1.0)
    o_hf.step()
    opt = B200Adam(mine.parameters(), lr=1e-3, betas=(0.9, 0.95))
    eng = TrainEngine(mine, opt, None, max_grad_norm=1.0, strategy="none", loss_scaler=scaler)
    assert eng.manual_optimization_step() is True
    sd2 = mine.state_dict()
    for n, p in hf.named_parameters():
        if n.endswith("query_key_value.bias"):
            continue  # its key third has an analytically zero gradient (softmax shift invariance): Adam turns that noise into +-lr
        upd_ref = p.detach().cpu() - sd[n]
        upd = sd2[n].cpu() - sd[n]
        assert rel(upd, upd_ref) <= 0.1, (n, rel(upd, upd_ref))  # first Adam step = lr * sign(g): sign flips of ~0 gradients only


def test_fp16_pythia_410m_layer_full_shape_vs_hf_fp32(dev):
    """The reference trains Pythia-410m in fp16 (src/models/pythia.py:33-41): one real layer + LM head, B 8 x 2049 tokens."""
    tr = pytest.importorskip("transformers")
    from multimodal_llm_pretraining_b200.models.configs import as_namespace, pythia_config_dict

    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        cfg = dict(pythia_config_dict("pythia-410m"), num_hidden_layers=1)
        torch.manual_seed(0)
        hf = tr.GPTNeoXForCausalLM(tr.GPTNeoXConfig(**cfg, attn_implementation="sdpa")).float()
        sd = {k: v.detach().clone() for k, v in hf.state_dict().items() if "inv_freq" not in k}
        hf = hf.to(dev).train()
        mine = B200GPTNeoXForCausalLM(as_namespace(cfg))
        mine.load_hf_state_dict(sd)
        mine.set_compute_dtype(FP16)
        mine = mine.to(dev).train()
        mine.loss_scale = torch.full((1,), 4096.0, device=dev)
        ids = torch.randint(0, cfg["vocab_size"], (8, 2049), generator=torch.Generator().manual_seed(2)).to(dev)
        loss = mine(input_ids=ids, labels=ids).loss
        loss.backward()
        ref = hf(input_ids=ids, labels=ids).loss
        ref.backward()
        assert abs(loss.item() - ref.item()) <= 2e-3 * ref.item()
        theirs = dict(hf.named_parameters())
        for n, p in mine.named_parameters():
            assert torch.isfinite(p.grad).all(), n
            e = rel(p.grad / 4096.0, theirs[n].grad)
            assert e <= 2e-2, (n, e)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


# ---------------------------------------------------------------------------------------------- device-side asserts
def _run_py(code: str):
    return subprocess.run([sys.executable, "-c", f"import sys; sys.path.insert(0, {str(ROOT)!r})\n" + code], capture_output=True, text=True, timeout=300)


@pytest.mark.skipif(not os.environ.get("B200_RUN_TRAP_TESTS"), reason="raises a deliberate device-side trap (an Xid on the host): opt-in with "
                    "B200_RUN_TRAP_TESTS=1; ran green on B200 twice this round (gpurun_out/r02_gputest1.log, r02_gputest4.log)")
def test_out_of_range_label_and_token_id_trap_like_torch():
    """A label outside [0, V) that is not ignore_index, or a token id outside the table, would read out of bounds: both trap
    with a message (torch raises a device-side assert). Own processes: a trap poisons the CUDA context. Opt-in: a trapped kernel is
    logged by the driver as a GPU exception, which a shared test box may treat as a fault of the lease."""
    r = _run_py("""
import torch
from multimodal_llm_pretraining_b200 import kernels as K
logits = torch.randn(8, 1024, device='cuda').to(torch.bfloat16)
labels = torch.randint(0, 1000, (8,), device='cuda'); labels[3] = 1003
try:
    K.cross_entropy_(logits, labels, V=1000); torch.cuda.synchronize(); print('NO ERROR')
except Exception as e:
    print('RAISED', type(e).__name__)
""")
    assert "NO ERROR" not in r.stdout and ("RAISED" in r.stdout or r.returncode != 0), r.stdout + r.stderr
    assert "outside [0, 1000)" in r.stdout + r.stderr
    r = _run_py("""
import torch
from multimodal_llm_pretraining_b200 import kernels as K
table = torch.randn(100, 64, device='cuda').to(torch.bfloat16)
ids = torch.tensor([1, 2, 100], device='cuda')
try:
    K.embedding_fwd(ids, table); torch.cuda.synchronize(); print('NO ERROR')
except Exception as e:
    print('RAISED', type(e).__name__)
""")
    assert "NO ERROR" not in r.stdout and ("RAISED" in r.stdout or r.returncode != 0), r.stdout + r.stderr
    assert "outside [0, 100)" in r.stdout + r.stderr


def test_adam_skip_flag_and_packed_gradients(dev):
    n = 65536 + 640
    p0 = torch.randn(2 * n, device=dev)
    g_full = torch.randn(2 * n, device=dev)
    own = (n, 2 * n)  # this "rank" owns the second half; its moments AND its gradients are packed
    cs = torch.tensor([own[0], own[0] + 65536], dtype=torch.int64, device=dev)
    cl = torch.tensor([65536, 640], dtype=torch.int32, device=dev)
    cg = torch.zeros(2, dtype=torch.int32, device=dev)
    cst = torch.tensor([0, 65536], dtype=torch.int64, device=dev)
    groups = [dict(lr=1e-2, beta1=0.9, beta2=0.95, eps=1e-8, weight_decay=0.0, bias_corr1=0.1, bias_corr2=0.05, adamw_mode=False)]
    res = []
    for packed in (False, True):
        p = p0.clone()
        m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
        sh = torch.zeros(2 * n, dtype=torch.bfloat16, device=dev)
        g = g_full[own[0]:own[1]].clone() if packed else g_full.clone()
        K.adam_step(p, g, m, v, sh, 0, cs, cl, cg, groups, chunk_state=cst, g_packed=packed, zero_grad=True)
        assert torch.equal(p[:n], p0[:n])
        assert float((g if packed else g[own[0]:own[1]]).abs().max()) == 0.0
        res.append((p, m, v, sh))
    for a, b in zip(*res):
        assert torch.equal(a, b)
    # skip flag: nothing moves, gradients still cleared
    p = p0.clone()
    m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    g = g_full.clone()
    flag = torch.ones(1, dtype=torch.int32, device=dev)
    K.adam_step(p, g, m, v, None, 0, cs, cl, cg, groups, chunk_state=cst, skip_flag=flag, zero_grad=True)
    assert torch.equal(p, p0) and float(m.abs().max()) == 0.0 and float(g[own[0]:own[1]].abs().max()) == 0.0
