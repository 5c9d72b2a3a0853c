"""One launch (reps = 1, for ncu) or a timed loop (reps > 1) of every GEMM class of the Pythia-1b step at its real shape, incl. the
LM head trio at T = 32768. usage: prof_gemm.py [reps]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from multimodal_llm_pretraining_b200 import kernels as K

dev = torch.device("cuda:0")
BF = torch.bfloat16
T, h, V = 32768, 2048, 50304
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
torch.manual_seed(0)


def run(tag, fn, flops):
    if reps == 1:
        fn()
        torch.cuda.synchronize()
        print(tag, flush=True)
        return
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{tag:40s} {ms:8.3f} ms {flops / ms / 1e9:8.1f} TFLOP/s", flush=True)


x = torch.randn(T, h, device=dev).to(BF)
res = torch.randn(T, h, device=dev).to(BF)
h4 = torch.randn(T, 4 * h, device=dev).to(BF)
pre = torch.empty(T, 4 * h, dtype=BF, device=dev)
q3 = torch.randn(T, 3 * h, device=dev).to(BF)
w_qkv = (torch.randn(3 * h, h, device=dev) * 0.02).to(BF)
w_d = (torch.randn(h, h, device=dev) * 0.02).to(BF)
w_up = (torch.randn(4 * h, h, device=dev) * 0.02).to(BF)
w_dn = (torch.randn(h, 4 * h, device=dev) * 0.02).to(BF)
b_qkv, b_d, b_up = torch.randn(3 * h, device=dev), torch.randn(h, device=dev), torch.randn(4 * h, device=dev)
F = 2.0 * T * h * h
run("fwd qkv +bias", lambda: K.gemm(x, w_qkv, bias=b_qkv), 3 * F)
run("fwd dense +bias+res", lambda: K.gemm(x, w_d, bias=b_d, residual=res), F)
run("fwd mlp_up +bias+gelu+aux", lambda: K.gemm(x, w_up, bias=b_up, gelu=True, aux_out=pre), 4 * F)
run("fwd mlp_down +bias+res", lambda: K.gemm(h4, w_dn, bias=b_d, residual=res), 4 * F)
run("dgrad mlp_down *dgelu", lambda: K.gemm(x, w_dn, b_mn=True, dgelu_in=pre), 4 * F)
run("dgrad mlp_up", lambda: K.gemm(h4, w_up, b_mn=True), 4 * F)
run("dgrad dense", lambda: K.gemm(x, w_d, b_mn=True), F)
run("dgrad qkv", lambda: K.gemm(q3, w_qkv, b_mn=True), 3 * F)
dw_up, dw_dn = torch.zeros(4 * h, h, device=dev), torch.zeros(h, 4 * h, device=dev)
dw_qkv, dw_d = torch.zeros(3 * h, h, device=dev), torch.zeros(h, h, device=dev)
run("wgrad mlp_up  [8192x2048]", lambda: K.gemm(h4, x, a_mn=True, b_mn=True, out=dw_up, accumulate=True), 4 * F)
run("wgrad mlp_down [2048x8192]", lambda: K.gemm(x, h4, a_mn=True, b_mn=True, out=dw_dn, accumulate=True), 4 * F)
run("wgrad qkv [6144x2048]", lambda: K.gemm(q3, x, a_mn=True, b_mn=True, out=dw_qkv, accumulate=True), 3 * F)
run("wgrad dense [2048x2048]", lambda: K.gemm(res, x, a_mn=True, b_mn=True, out=dw_d, accumulate=True), F)
del h4, pre, q3, dw_up, dw_dn
w_out = (torch.randn(V, h, device=dev) * 0.02).to(BF)
logits = torch.empty(T, V, dtype=BF, device=dev)
FL = 2.0 * T * h * V
run("fwd lm_head", lambda: K.gemm(x, w_out, out=logits), FL)
logits.normal_(0, 1e-3)
dw_out = torch.zeros(V, h, device=dev)
run("wgrad lm_head", lambda: K.gemm(logits, x, a_mn=True, b_mn=True, out=dw_out, accumulate=True), FL)
run("dgrad lm_head", lambda: K.gemm(logits, w_out, b_mn=True), FL)
print("ok")
