"""Is the head_dim-256 attention bound by a per-SM resource or by a chip-wide one (L2 slices / HBM)? Times the forward and the
backward with fewer CTAs than SMs up to the full Pythia-1b shape: if ns per (CTA x key block) grows with the number of co-running
CTAs, the bound is shared. usage: attn_scaling.py [D]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from multimodal_llm_pretraining_b200 import kernels as K

dev = torch.device("cuda:0")
D = int(sys.argv[1]) if len(sys.argv) > 1 else 256
S = 2048


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


print(f"head_dim {D}, S {S}, causal; 16 query tiles of 128 per (b, h); critical path = the last query tile = {S // 64} key blocks of 64")
print(f"{'B':>3s} {'H':>3s} {'CTAs':>6s} {'fwd ms':>8s} {'fwd TF/s':>9s} {'bwd ms':>8s} {'bwd TF/s':>9s}")
for B, H in [(1, 1), (1, 4), (1, 8), (2, 8), (4, 8), (16, 8)]:
    qkv = (torch.randn(B, S, H, 3, D, device=dev) * 0.5).to(torch.bfloat16)
    q, k, v = qkv[:, :, :, 0], qkv[:, :, :, 1], qkv[:, :, :, 2]
    o, lse = K.attention_fwd(q, k, v, causal=True)
    do = torch.randn_like(o)
    dqkv = torch.empty_like(qkv)
    f = timeit(lambda: K.attention_fwd(q, k, v, causal=True))
    b = timeit(lambda: K.attention_bwd(q, k, v, o, lse, do, dqkv[:, :, :, 0], dqkv[:, :, :, 1], dqkv[:, :, :, 2], causal=True))
    fl = 4.0 * B * H * S * S * D / 2
    print(f"{B:3d} {H:3d} {B * H * 16:6d} {f:8.4f} {fl / f / 1e9:9.1f} {b:8.4f} {2.5 * fl / b / 1e9:9.1f}", flush=True)
    del qkv, o, lse, do, dqkv
    K._score_cache.clear()
