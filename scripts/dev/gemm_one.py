"""Runs a few launches of selected GEMM variants (for ncu captures). usage: gemm_one.py [gelu] [res] [plain] [wgrad] [dgelu]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from multimodal_llm_pretraining_b200 import kernels as K

dev = torch.device("cuda:0")
T, h = 32768, 2048
BF = torch.bfloat16
which = sys.argv[1:] or ["gelu"]
x = torch.randn(T, h, device=dev).to(BF)
w = (torch.randn(4 * h, h, device=dev) * 0.02).to(BF)
bias = torch.randn(4 * h, device=dev)
res = torch.randn(T, 4 * h, device=dev).to(BF)
aux = torch.empty(T, 4 * h, device=dev, dtype=BF)
dy = torch.randn(T, 4 * h, device=dev).to(BF)
dw = torch.zeros(4 * h, h, device=dev)
for _ in range(4):
    if "gelu" in which:
        K.gemm(x, w, bias=bias, gelu=True, aux_out=aux)
    if "res" in which:
        K.gemm(x, w, bias=bias, residual=res)
    if "plain" in which:
        K.gemm(x, w, bias=bias)
    if "wgrad" in which:
        K.gemm(dy, x, a_mn=True, b_mn=True, out=dw, accumulate=True)
    if "dgelu" in which:
        K.gemm(x, w.t().contiguous(), b_mn=True, dgelu_in=res)
torch.cuda.synchronize()
print("ok")
