"""tcgen05 flash attention (C ABI) vs an fp32 PyTorch statement of softmax(QK^T * scale [+causal]) V, outputs and
gradients. Tolerance rel <= 2e-2 (bf16 operands, fp32 accumulate) as north_star states."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from multimodal_llm_pretraining_b200 import kernels as K  # noqa: E402

BF16 = torch.bfloat16


def rel_err(got, ref):
    got, ref = got.float(), ref.float()
    return ((got - ref).norm() / (ref.norm() + 1e-12)).item()


def ref_attention(q, k, v, causal, scale):
    # q,k,v fp32 [B,S,H,D]
    qh, kh, vh = (t.permute(0, 2, 1, 3) for t in (q, k, v))
    s = (qh @ kh.transpose(-1, -2)) * scale
    if causal:
        S = q.shape[1]
        mask = torch.ones(S, S, dtype=torch.bool, device=q.device).tril()
        s = s.masked_fill(~mask, float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = (p @ vh).permute(0, 2, 1, 3)
    lse = torch.logsumexp(s, dim=-1)  # [B,H,S]
    return o, lse


CASES = [
    # B, S, H, D, causal, packed
    (2, 256, 2, 64, True, True),
    (1, 512, 3, 64, False, True),
    (2, 384, 2, 128, True, True),
    (1, 256, 2, 128, False, False),
    (2, 256, 2, 256, True, True),
    (1, 640, 1, 256, True, True),
    (1, 384, 2, 256, False, False),
    (1, 200, 2, 64, True, True),     # ragged: S not a multiple of the tile
    (1, 136, 1, 256, True, True),
    (1, 2048, 2, 256, True, True),   # Pythia-1b head shape at full sequence length
    (1, 2048, 2, 64, True, True),
    (2, 512, 2, 256, False, True),   # head_dim 256, bidirectional, score-scratch path
    (2, 768, 3, 256, True, True),    # three 256-key tiles per head
    (2, 256, 4, 80, True, True),     # Pythia-2.8b head_dim 80: zero-padded to 128 by 3-D tensor maps
    (1, 200, 2, 80, True, True),
    (1, 384, 2, 80, False, False),
]


def make_qkv(B, S, H, D, packed, dev, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    if packed:  # GPT-NeoX layout [B,S,H,3,D]
        buf = torch.randn(B, S, H, 3, D, generator=g).to(dev).to(BF16)
        return buf, buf[:, :, :, 0], buf[:, :, :, 1], buf[:, :, :, 2]
    q = torch.randn(B, S, H, D, generator=g).to(dev).to(BF16)
    k = torch.randn(B, S, H, D, generator=g).to(dev).to(BF16)
    v = torch.randn(B, S, H, D, generator=g).to(dev).to(BF16)
    return None, q, k, v


@pytest.mark.parametrize("B,S,H,D,causal,packed", CASES)
def test_attention_fwd_bwd(dev, B, S, H, D, causal, packed):
    _, q, k, v = make_qkv(B, S, H, D, packed, dev, seed=S + D)
    scale = D ** -0.5
    o, lse = K.attention_fwd(q, k, v, causal, scale)
    qf, kf, vf = (t.float().detach().clone().requires_grad_(True) for t in (q, k, v))
    ro, rlse = ref_attention(qf, kf, vf, causal, scale)
    assert torch.isfinite(o.float()).all()
    e = rel_err(o, ro)
    assert e <= 2e-2, f"O rel err {e:.3e}"
    e = (lse - rlse).abs().max().item()
    assert e <= 2e-2, f"LSE max abs err {e:.3e}"

    g = torch.Generator(device="cpu").manual_seed(7)
    d_o = torch.randn(B, S, H, D, generator=g).to(dev).to(BF16)
    ro.backward(d_o.float())
    if packed:
        dbuf = torch.full((B, S, H, 3, D), float("nan"), device=dev, dtype=BF16)
        dq, dk, dv = dbuf[:, :, :, 0], dbuf[:, :, :, 1], dbuf[:, :, :, 2]
    else:
        dq, dk, dv = (torch.full((B, S, H, D), float("nan"), device=dev, dtype=BF16) for _ in range(3))
    K.attention_bwd(q, k, v, o, lse, d_o, dq, dk, dv, causal, scale)
    for name, got, ref in (("dQ", dq, qf.grad), ("dK", dk, kf.grad), ("dV", dv, vf.grad)):
        assert torch.isfinite(got.float()).all(), f"{name} has non-finite values"
        e = rel_err(got, ref)
        assert e <= 2e-2, f"{name} rel err {e:.3e}"


def test_attention_large_logits(dev):
    # running-max path: growing scores force O rescales
    B, S, H, D = 1, 512, 1, 64
    g = torch.Generator(device="cpu").manual_seed(3)
    q = (torch.randn(B, S, H, D, generator=g) * 4).to(dev).to(BF16)
    k = (torch.randn(B, S, H, D, generator=g) * 4).to(dev).to(BF16)
    k = (k.float() * torch.linspace(0.1, 3.0, S, device=dev)[None, :, None, None]).to(BF16)
    v = torch.randn(B, S, H, D, generator=g).to(dev).to(BF16)
    o, lse = K.attention_fwd(q, k, v, False, D ** -0.5)
    ro, rlse = ref_attention(q.float(), k.float(), v.float(), False, D ** -0.5)
    assert rel_err(o, ro) <= 2e-2
    assert ((lse - rlse).abs() / rlse.abs().clamp_min(1)).max().item() <= 1e-2


@pytest.mark.parametrize("use_scratch", [True, False])
def test_attention_bwd_d256_scratch_paths_agree_and_ignore_stale_scratch(dev, use_scratch):
    """head_dim 256: dK/dV via materialised P/dS + batched GEMMs (scratch poisoned with NaN first: every element the
    GEMMs read must have been written by this call) vs the fused recompute pass."""
    B, S, H, D = 2, 1024, 2, 256
    _, q, k, v = make_qkv(B, S, H, D, True, dev, seed=11)
    scale = D ** -0.5
    o, lse = K.attention_fwd(q, k, v, True, scale)
    qf, kf, vf = (t.float().detach().clone().requires_grad_(True) for t in (q, k, v))
    ro, _ = ref_attention(qf, kf, vf, True, scale)
    d_o = torch.randn(B, S, H, D, generator=torch.Generator(device="cpu").manual_seed(5)).to(dev).to(BF16)
    ro.backward(d_o.float())
    old = K.USE_SCORE_SCRATCH
    K.USE_SCORE_SCRATCH = use_scratch
    try:
        if use_scratch:
            ps, dss = K._score_scratch(q.device, B * H, S)
            ps.fill_(float("nan")), dss.fill_(float("nan"))
        dbuf = torch.full((B, S, H, 3, D), float("nan"), device=dev, dtype=BF16)
        dq, dk, dv = dbuf[:, :, :, 0], dbuf[:, :, :, 1], dbuf[:, :, :, 2]
        K.attention_bwd(q, k, v, o, lse, d_o, dq, dk, dv, True, scale)
    finally:
        K.USE_SCORE_SCRATCH = old
    for name, got, ref in (("dQ", dq, qf.grad), ("dK", dk, kf.grad), ("dV", dv, vf.grad)):
        assert torch.isfinite(got.float()).all(), f"{name} has non-finite values"
        assert rel_err(got, ref) <= 2e-2, f"{name} rel err {rel_err(got, ref):.3e}"


def _drop_keep_mask(seed, B, H, S, p):
    """Host restatement of the kernels' counter-based mask (attention.cu attn_drop_key / attn_drop_row / attn_drop_hash /
    attn_drop_keep): one 32-bit hash per (query, key pair), 16-bit lane k & 1 compared with round(p * 65536)."""
    import numpy as np

    thr = np.uint32(int(p * 65536.0 + 0.5))
    M = np.uint64(0xFFFFFFFF)
    u = lambda v: np.uint64(v) & M  # noqa: E731 - 32-bit wrap-around arithmetic carried in uint64
    s0 = u(u(seed & 0xFFFFFFFF) * np.uint64(0x9E3779B1) + np.uint64(0x85EBCA6B))
    s1 = u((seed >> 32) & 0xFFFFFFFF) ^ np.uint64(0xC2B2AE35)
    half = np.uint64((S + 1) // 2)
    bh = np.arange(B * H, dtype=np.uint64)[:, None, None]
    q = np.arange(S, dtype=np.uint64)[None, :, None]
    k = np.arange(S, dtype=np.uint64)[None, None, :]
    row = ((bh * np.uint64(S) + q) & M) * half & M
    x = ((row + (k >> np.uint64(1))) & M) ^ s0
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & M
    x ^= x >> np.uint64(15)
    x = (x + s1) & M
    x = (x * np.uint64(0x846CA68B)) & M
    x ^= x >> np.uint64(16)
    bits = (x >> (np.uint64(16) * (k & np.uint64(1)))) & np.uint64(0xFFFF)
    keep = bits >= np.uint64(thr)
    return torch.from_numpy(keep.reshape(B, H, S, S)), 65536.0 / (65536.0 - float(thr))


@pytest.mark.parametrize("B,S,H,D,causal", [(2, 256, 2, 64, False), (1, 384, 2, 64, True), (1, 256, 2, 128, False), (1, 256, 1, 256, True)])
def test_attention_dropout_matches_reference_with_same_mask(dev, B, S, H, D, causal):
    """Softmax-probability dropout: forward output and all three gradients against an fp32 reference that applies the SAME
    mask (reconstructed on the host from the counter-based rule); keep rate statistics."""
    p_drop, seed = 0.1, 0x1234ABCD5678
    _, q, k, v = make_qkv(B, S, H, D, True, dev, seed=S + D + 1)
    scale = D ** -0.5
    keep, inv_keep = _drop_keep_mask(seed, B, H, S, p_drop)
    assert abs(keep.float().mean().item() - (1 - p_drop)) < 5e-3
    keep = keep.to(dev)
    qf, kf, vf = (t.float().detach().clone().requires_grad_(True) for t in (q, k, v))
    qh, kh, vh = (t.permute(0, 2, 1, 3) for t in (qf, kf, vf))
    sc = (qh @ kh.transpose(-1, -2)) * scale
    if causal:
        sc = sc.masked_fill(~torch.ones(S, S, dtype=torch.bool, device=dev).tril(), float("-inf"))
    pr = torch.softmax(sc, dim=-1) * keep * inv_keep
    ro = (pr @ vh).permute(0, 2, 1, 3)
    o, lse = K.attention_fwd(q, k, v, causal, scale, dropout_p=p_drop, dropout_seed=seed)
    assert rel_err(o, ro) <= 2e-2, rel_err(o, ro)
    d_o = torch.randn(B, S, H, D, generator=torch.Generator(device="cpu").manual_seed(9)).to(dev).to(BF16)
    ro.backward(d_o.float())
    dbuf = torch.full((B, S, H, 3, D), float("nan"), device=dev, dtype=BF16)
    dq, dk, dv = dbuf[:, :, :, 0], dbuf[:, :, :, 1], dbuf[:, :, :, 2]
    K.attention_bwd(q, k, v, o, lse, d_o, dq, dk, dv, causal, scale, dropout_p=p_drop, dropout_seed=seed)
    for name, got, ref in (("dQ", dq, qf.grad), ("dK", dk, kf.grad), ("dV", dv, vf.grad)):
        assert torch.isfinite(got.float()).all(), name
        assert rel_err(got, ref) <= 2e-2, (name, rel_err(got, ref))
