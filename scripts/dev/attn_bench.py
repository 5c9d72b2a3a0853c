"""Times flash attention fwd / bwd on the benchmarked head shapes (packed GPT-NeoX qkv layout)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from multimodal_llm_pretraining_b200 import kernels as K

dev = torch.device("cuda:0")
BF = torch.bfloat16


def timeit(fn, iters=10):
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


def main():
    cases = [("pythia-1b", 16, 2048, 8, 256, True), ("1b-noncausal", 16, 2048, 8, 256, False), ("pythia-410m", 16, 2048, 16, 64, True), ("pythia-1.4b", 8, 2048, 16, 128, True),
             ("roberta-large", 64, 512, 16, 64, False)]
    only = sys.argv[1:] or None
    print(f"{'case':16s} {'fwd ms':>8s} {'fwd TF/s':>9s} {'bwd ms':>8s} {'bwd TF/s':>9s}   (FLOPs = useful: causal halves)")
    for name, B, S, H, D, causal in cases:
        if only and name not in only:
            continue
        qkv = torch.randn(B, S, H, 3, D, device=dev).to(BF)
        q, k, v = qkv[:, :, :, 0], qkv[:, :, :, 1], qkv[:, :, :, 2]
        o, lse = K.attention_fwd(q, k, v, causal=causal)
        d_o = torch.randn_like(o)
        dqkv = torch.empty_like(qkv)
        dq, dk, dv = dqkv[:, :, :, 0], dqkv[:, :, :, 1], dqkv[:, :, :, 2]
        f = 4.0 * B * H * S * S * D * (0.5 if causal else 1.0)
        tf = timeit(lambda: K.attention_fwd(q, k, v, causal=causal))
        tb = timeit(lambda: K.attention_bwd(q, k, v, o, lse, d_o, dq, dk, dv, causal=causal))
        print(f"{name:16s} {tf:8.3f} {f / tf / 1e9:9.1f} {tb:8.3f} {2.5 * f / tb / 1e9:9.1f}")


if __name__ == "__main__":
    main()
