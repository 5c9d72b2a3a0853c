"""One tiny Pythia-1b-shaped workload for ncu captures: a plain fwd GEMM (mlp_up shape), a wgrad GEMM, attention fwd + bwd
(head_dim 256), cross entropy, LayerNorm, Adam chunk. usage: prof_step.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from multimodal_llm_pretraining_b200 import kernels as K

dev = torch.device("cuda:0")
BF = torch.bfloat16
T, h, V = 32768, 2048, 50304
x = torch.randn(T, h, device=dev).to(BF)
w = (torch.randn(4 * h, h, device=dev) * 0.02).to(BF)
bias = torch.randn(4 * h, device=dev)
dy = torch.randn(T, 4 * h, device=dev).to(BF)
dw = torch.zeros(4 * h, h, device=dev)
qkv = torch.randn(16, 2048, 8, 3, 256, device=dev).to(BF)
q, k, v = qkv[:, :, :, 0], qkv[:, :, :, 1], qkv[:, :, :, 2]
dqkv = torch.empty_like(qkv)
logits = torch.randn(8192, V, device=dev).to(BF)
labels = torch.randint(0, V, (8192,), device=dev)
g1, b1, g2, b2 = (torch.randn(h, device=dev) for _ in range(4))
for _ in range(2):
    K.gemm(x, w, bias=bias)
    K.gemm(dy, x, a_mn=True, b_mn=True, out=dw, accumulate=True)
    o, lse = K.attention_fwd(q, k, v, causal=True)
    K.attention_bwd(q, k, v, o, lse, o, dqkv[:, :, :, 0], dqkv[:, :, :, 1], dqkv[:, :, :, 2], causal=True)
    K.cross_entropy_(logits, labels, V=V, write_grad=True)
    y1, y2, mean, rstd = K.layernorm_fwd(x, g1, b1, 1e-5, g2, b2)
torch.cuda.synchronize()
print("ok")
