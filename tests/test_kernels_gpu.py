"""Parity of every libb200pt kernel (through the C ABI) against a plain fp32 PyTorch statement of the same op.

Tolerances: bf16 outputs are compared with the fp32 reference at rel <= 2e-2 of the reference's RMS (north_star:
"per-kernel outputs and gradients within bf16 tolerance, rel <= 2e-2 vs an fp32 reference"); fp32 outputs tighter.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from multimodal_llm_pretraining_b200 import kernels as K  # noqa: E402

BF16, F32 = torch.bfloat16, torch.float32


def rel_err(got, ref):
    got, ref = got.float(), ref.float()
    return ((got - ref).norm() / (ref.norm() + 1e-12)).item()


def max_err(got, ref):
    return (got.float() - ref.float()).abs().max().item()


def check(got, ref, tol, name):
    assert torch.isfinite(got.float()).all(), f"{name}: non-finite output"
    r = rel_err(got, ref)
    assert r <= tol, f"{name}: rel err {r:.3e} > {tol:.1e} (max abs {max_err(got, ref):.3e}, ref rms {ref.float().pow(2).mean().sqrt().item():.3e})"


# ---------------------------------------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize("rows,cols", [(64, 768), (257, 1024), (128, 2048), (96, 2560), (33, 4096), (40, 5120), (16, 128)])
@pytest.mark.parametrize("dual", [False, True])
def test_layernorm_fwd_bwd(dev, rows, cols, dual):
    g = torch.Generator(device="cpu").manual_seed(rows * 7 + cols)
    x = (torch.randn(rows, cols, generator=g) * 1.5 + 0.3).to(dev).to(BF16)
    g1 = (1 + 0.1 * torch.randn(cols, generator=g)).to(dev)
    b1 = (0.1 * torch.randn(cols, generator=g)).to(dev)
    g2 = (1 + 0.1 * torch.randn(cols, generator=g)).to(dev) if dual else None
    b2 = (0.1 * torch.randn(cols, generator=g)).to(dev) if dual else None
    dy1 = torch.randn(rows, cols, generator=g).to(dev).to(BF16)
    dy2 = torch.randn(rows, cols, generator=g).to(dev).to(BF16) if dual else None
    dres = torch.randn(rows, cols, generator=g).to(dev).to(BF16)

    y1, y2, mean, rstd = K.layernorm_fwd(x, g1, b1, 1e-5, g2, b2)
    xf = x.float().requires_grad_(True)
    g1r, b1r = g1.clone().requires_grad_(True), b1.clone().requires_grad_(True)
    r1 = torch.nn.functional.layer_norm(xf, (cols,), g1r, b1r, 1e-5)
    check(y1, r1, 1e-2, "ln y1")
    check(mean, xf.mean(1), 1e-4, "ln mean")
    loss = (r1 * dy1.float()).sum() + (xf * dres.float()).sum()
    if dual:
        g2r, b2r = g2.clone().requires_grad_(True), b2.clone().requires_grad_(True)
        r2 = torch.nn.functional.layer_norm(xf, (cols,), g2r, b2r, 1e-5)
        check(y2, r2, 1e-2, "ln y2")
        loss = loss + (r2 * dy2.float()).sum()
    loss.backward()

    dg1 = torch.full((cols,), 0.5, device=dev)  # accumulate semantics: starts non-zero
    db1 = torch.full((cols,), -0.25, device=dev)
    dg2 = torch.zeros(cols, device=dev) if dual else None
    db2 = torch.zeros(cols, device=dev) if dual else None
    dx = K.layernorm_bwd(x, mean, rstd, g1, dy1, dg1, db1, g2, dy2, dg2, db2, dres=dres)
    check(dx, xf.grad, 1e-2, "ln dx")
    check(dg1 - 0.5, g1r.grad, 2e-3, "ln dgamma")
    check(db1 + 0.25, b1r.grad, 2e-3, "ln dbeta")
    if dual:
        check(dg2, g2r.grad, 2e-3, "ln dgamma2")
        check(db2, b2r.grad, 2e-3, "ln dbeta2")


# ---------------------------------------------------------------------------------------------- GELU / RoPE / embedding
def test_gelu(dev):
    x = (torch.randn(4096, 1024, device=dev) * 2).to(BF16)
    dy = torch.randn(4096, 1024, device=dev).to(BF16)
    xf = x.float().requires_grad_(True)
    ref = torch.nn.functional.gelu(xf)
    ref.backward(dy.float())
    check(K.gelu_fwd(x), ref, 5e-3, "gelu fwd")
    check(K.gelu_bwd(x, dy), xf.grad, 5e-3, "gelu bwd")


def _rope_ref(qkv, cos, sin, B, S, nh, hd, rot):
    # HF apply_rotary_pos_emb / rotate_half (modeling_gpt_neox.py:119-159) in fp32
    x = qkv.float().view(B, S, nh, 3, hd).clone()
    c = torch.cat([cos, cos], -1)[None, :S, None, :]
    s = torch.cat([sin, sin], -1)[None, :S, None, :]
    for w in (0, 1):
        t = x[:, :, :, w, :rot]
        rh = torch.cat([-t[..., rot // 2:], t[..., : rot // 2]], -1)
        x[:, :, :, w, :rot] = t * c + rh * s
    return x.view(B * S, nh * 3 * hd)


@pytest.mark.parametrize("nh,hd,rot", [(8, 256, 64), (16, 64, 16), (4, 80, 20), (4, 128, 32)])
def test_rope(dev, nh, hd, rot):
    B, S = 2, 96
    qkv = torch.randn(B * S, nh * 3 * hd, device=dev).to(BF16)
    inv = 1.0 / (10000 ** (torch.arange(0, rot, 2, device=dev).float() / rot))
    ang = torch.arange(S, device=dev).float()[:, None] * inv[None, :]
    cos, sin = ang.cos().contiguous(), ang.sin().contiguous()
    ref = _rope_ref(qkv, cos, sin, B, S, nh, hd, rot)
    got = K.rope_qk_inplace(qkv.clone(), cos, sin, B, S, nh, hd, rot)
    check(got, ref, 5e-3, "rope fwd")
    back = K.rope_qk_inplace(got.clone(), cos, sin, B, S, nh, hd, rot, inverse=True)
    check(back, qkv, 1e-2, "rope inverse round trip")


def test_embedding(dev):
    V, h, T = 1000, 768, 4096
    table = torch.randn(V, h, device=dev).to(BF16)
    ids = torch.randint(0, V, (T,), device=dev)
    out = K.embedding_fwd(ids, table)
    assert torch.equal(out, table[ids]), "embedding gather must be bit exact"
    dout = torch.randn(T, h, device=dev).to(BF16)
    dtab = torch.zeros(V, h, device=dev)
    K.embedding_bwd(ids, dout, dtab)
    ref = torch.zeros(V, h, device=dev).index_add_(0, ids, dout.float())
    check(dtab, ref, 1e-5, "embedding bwd")
    t1 = torch.randn(600, h, device=dev).to(BF16)
    ids1 = torch.randint(0, 600, (T,), device=dev)
    t2 = torch.randn(1, h, device=dev).to(BF16)
    ids2 = torch.zeros(T, dtype=torch.int64, device=dev)
    out3 = K.embedding3_fwd(ids, table, ids1, t1, ids2, t2)
    check(out3, table[ids].float() + t1[ids1].float() + t2[ids2].float(), 5e-3, "embedding3")


# ---------------------------------------------------------------------------------------------- cross entropy
@pytest.mark.parametrize("T,V,ld", [(64, 50304, 50304), (37, 50265, 50304), (128, 1000, 1000), (16, 65536, 65536)])
def test_cross_entropy(dev, T, V, ld):
    g = torch.Generator(device="cpu").manual_seed(V)
    buf = (torch.randn(T, ld, generator=g) * 3).to(dev).to(BF16)
    labels = torch.randint(0, V, (T,), generator=g).to(dev)
    labels[::5] = -100
    xf = buf[:, :V].float().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(xf, labels, ignore_index=-100)
    ref.backward()
    logits = buf.clone()
    loss, n_valid = K.cross_entropy_(logits, labels, V=V)
    assert n_valid.item() == (labels != -100).sum().item()
    assert abs(loss.item() - ref.item()) <= 1e-4 * abs(ref.item()), f"loss {loss.item()} vs {ref.item()}"
    check(logits[:, :V], xf.grad, 1e-2, "dlogits")
    if ld > V:
        assert (logits[:, V:] == 0).all(), "padding columns of dlogits must be zero"
    # loss-only mode leaves the logits untouched
    logits2 = buf.clone()
    loss2, _ = K.cross_entropy_(logits2, labels, V=V, write_grad=False)
    assert torch.equal(logits2, buf) and abs(loss2.item() - loss.item()) < 1e-6


# ---------------------------------------------------------------------------------------------- GEMM
def _gemm_ref(A, B, a_mn, b_mn):
    Af = A.float().t() if a_mn else A.float()
    Bf = B.float() if b_mn else B.float().t()
    return Af @ Bf


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 256), (384, 768, 1024), (200, 136, 328), (128, 128, 128), (4096, 2048, 2048)])
def test_gemm_layouts(dev, a_mn, b_mn, M, N, K):
    g = torch.Generator(device="cpu").manual_seed(M + 3 * N + 7 * K)
    A = torch.randn((K, M) if a_mn else (M, K), generator=g).to(dev).to(BF16)
    B = torch.randn((K, N) if b_mn else (N, K), generator=g).to(dev).to(BF16)
    ref = _gemm_ref(A, B, a_mn, b_mn)
    got = K_gemm(A, B, a_mn=a_mn, b_mn=b_mn)
    check(got, ref, 5e-3, f"gemm a_mn={a_mn} b_mn={b_mn} {M}x{N}x{K}")
    got32 = K_gemm(A, B, a_mn=a_mn, b_mn=b_mn, out_dtype=F32)
    check(got32, ref, 1e-4, f"gemm fp32 out a_mn={a_mn} b_mn={b_mn} {M}x{N}x{K}")


def K_gemm(*a, **k):
    return K.gemm(*a, **k)


@pytest.mark.parametrize("M,N,Kd", [(512, 1024, 512), (200, 136, 328), (128, 256, 64), (300, 520, 192)])
def test_gemm_epilogues(dev, M, N, Kd):
    g = torch.Generator(device="cpu").manual_seed(5)
    A = torch.randn(M, Kd, generator=g).to(dev).to(BF16)
    W = (torch.randn(N, Kd, generator=g) * 0.05).to(dev).to(BF16)
    bias = torch.randn(N, generator=g).to(dev)
    res = torch.randn(M, N, generator=g).to(dev).to(BF16)
    base = A.float() @ W.float().t()
    check(K.gemm(A, W, bias=bias), base + bias, 5e-3, "bias")
    aux = torch.empty(M, N, dtype=BF16, device=dev)
    out = K.gemm(A, W, bias=bias, gelu=True, aux_out=aux)
    check(aux, base + bias, 5e-3, "aux pre-activation")
    check(out, torch.nn.functional.gelu(base + bias), 5e-3, "bias+gelu")
    check(K.gemm(A, W, bias=bias, residual=res), base + bias + res.float(), 5e-3, "bias+residual")
    alpha = torch.tensor(0.125, device=dev)
    check(K.gemm(A, W, alpha=alpha), base * 0.125, 5e-3, "alpha")
    h = torch.randn(M, N, generator=g).to(dev).to(BF16)
    hf = h.float().requires_grad_(True)
    torch.nn.functional.gelu(hf).backward(base)
    Wt = W.t().contiguous()  # dgrad layout: B stored [K, N]
    check(K.gemm(A, Wt, b_mn=True, dgelu_in=h), hf.grad, 5e-3, "fused dgelu")
    cs = torch.full((N,), 0.25, device=dev)  # accumulate semantics: starts non-zero
    check(K.gemm(A, Wt, b_mn=True, dgelu_in=h, colsum_out=cs), hf.grad, 5e-3, "fused dgelu + column sums (output)")
    check(cs - 0.25, hf.grad.sum(0), 2e-3, "column sums of the dgelu dgrad (bias gradient) from the epilogue")
    acc = torch.randn(M, N, generator=g).to(dev)
    want = acc + base
    K.gemm(A, W, out=acc, accumulate=True)
    check(acc, want, 1e-4, "fp32 accumulate")
    accb = res.clone()
    K.gemm(A, W, out=accb, accumulate=True)
    check(accb, res.float() + base, 5e-3, "bf16 accumulate")


def test_gemm_wgrad_shape(dev):
    # dW[out,in] += dY^T X with a long reduction (tokens) — the wgrad configuration
    T, out_f, in_f = 8192, 768, 512
    g = torch.Generator(device="cpu").manual_seed(11)
    dY = torch.randn(T, out_f, generator=g).to(dev).to(BF16)
    X = torch.randn(T, in_f, generator=g).to(dev).to(BF16)
    dW = torch.zeros(out_f, in_f, device=dev)
    K.gemm(dY, X, a_mn=True, b_mn=True, out=dW, accumulate=True)
    K.gemm(dY, X, a_mn=True, b_mn=True, out=dW, accumulate=True)
    check(dW, 2 * (dY.float().t() @ X.float()), 1e-4, "wgrad accumulate x2")


# ---------------------------------------------------------------------------------------------- optimizer
@pytest.mark.parametrize("adamw", [False, True])
def test_adam_matches_torch(dev, adamw):
    torch.manual_seed(0)
    sizes = [(1024, 64), (777,), (65536 * 2 + 13,), (5, 5)]
    n_total, offs = 0, []
    for s in sizes:
        offs.append(n_total)
        n_total += (math.prod(s) + 63) // 64 * 64
    flat_p = torch.zeros(n_total, device=dev)
    flat_g = torch.zeros(n_total, device=dev)
    m = torch.zeros(n_total, device=dev)
    v = torch.zeros(n_total, device=dev)
    shadow = torch.zeros(n_total, device=dev, dtype=BF16)
    ref_params = []
    for s, o in zip(sizes, offs):
        n = math.prod(s)
        flat_p[o:o + n] = torch.randn(n, device=dev)
        ref_params.append(flat_p[o:o + n].clone().view(s).requires_grad_(True))
    wds = [0.01, 0.0, 0.01, 0.0]
    cls = torch.optim.AdamW if adamw else torch.optim.Adam
    opt = cls([{"params": [p], "weight_decay": wd} for p, wd in zip(ref_params, wds)], lr=3e-3, betas=(0.9, 0.95), eps=1e-8)
    starts, lens, grps = [], [], []
    for i, (s, o) in enumerate(zip(sizes, offs)):
        n = math.prod(s)
        for c in range(0, n, 65536):
            starts.append(o + c)
            lens.append(min(65536, n - c))
            grps.append(0 if wds[i] > 0 else 1)
    cs = torch.tensor(starts, dtype=torch.int64, device=dev)
    cl = torch.tensor(lens, dtype=torch.int32, device=dev)
    cg = torch.tensor(grps, dtype=torch.int32, device=dev)
    for step in range(1, 6):
        for p, s, o in zip(ref_params, sizes, offs):
            gr = torch.randn(s, device=dev)
            p.grad = gr.clone()
            flat_g[o:o + gr.numel()] = gr.flatten()
        opt.step()
        groups = [dict(lr=3e-3, beta1=0.9, beta2=0.95, eps=1e-8, weight_decay=wd, bias_corr1=1 - 0.9 ** step,
                       bias_corr2=1 - 0.95 ** step, adamw_mode=adamw) for wd in (0.01, 0.0)]
        K.adam_step(flat_p, flat_g, m, v, shadow, 0, cs, cl, cg, groups, zero_grad=True)
    assert (flat_g == 0).all(), "zero_grad pass must clear the gradients"
    for p, s, o in zip(ref_params, sizes, offs):
        n = math.prod(s)
        check(flat_p[o:o + n], p.detach().flatten(), 2e-6, "adam params")
        assert torch.equal(shadow[o:o + n], flat_p[o:o + n].to(BF16)), "bf16 shadow must be the rounded fp32 master"


def test_sumsq_clip_cast(dev):
    x = torch.randn(1_000_003, device=dev)
    out = torch.zeros((), device=dev)
    K.sumsq_(x[:1_000_000], out)
    K.sumsq_(x[:1_000_000], out)
    ref = 2 * x[:1_000_000].double().pow(2).sum().item()
    assert abs(out.item() - ref) <= 1e-5 * ref
    norm, coef = K.clip_coef(out, 1.0)
    assert abs(norm.item() - math.sqrt(ref)) <= 1e-5 * math.sqrt(ref)
    assert abs(coef.item() - 1.0 / (math.sqrt(ref) + 1e-6)) <= 1e-6
    _, coef0 = K.clip_coef(out, 0.0)
    assert coef0.item() == 1.0
    y = torch.empty(1_000_003, dtype=BF16, device=dev)
    K.cast_f32_to_bf16(x, y)
    assert torch.equal(y, x.to(BF16))
