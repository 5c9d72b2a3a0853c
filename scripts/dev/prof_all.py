"""One launch of EVERY kernel class of the hot path at the Pythia-1b step shapes (plus the head_dim-64 attention kernels of
Pythia-410m / RoBERTa), for `ncu --set full` captures (profiles/rNN_ncu_all_kernels.txt). Each op prints its tag so that the
launch order in the report can be mapped back. usage: prof_all.py [reps]   (reps > 1 also prints CUDA-event timings)"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from multimodal_llm_pretraining_b200 import kernels as K
from multimodal_llm_pretraining_b200.modeling_gpt_neox import B200GPTNeoXForCausalLM
from multimodal_llm_pretraining_b200.models.configs import as_namespace, pythia_config_dict
from multimodal_llm_pretraining_b200.optim import B200Adam

dev = torch.device("cuda:0")
BF = torch.bfloat16
T, h, V, B, S, nh, hd = 32768, 2048, 50304, 16, 2048, 8, 256
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
torch.manual_seed(0)


def run(tag, fn, nbytes=None, flops=None):
    if reps == 1:
        fn()
        torch.cuda.synchronize()
        print(tag, flush=True)
        return
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    extra = ""
    if nbytes:
        extra += f"  {nbytes / ms / 1e6:8.1f} GB/s (algorithmic {nbytes / 1e9:.3f} GB)"
    if flops:
        extra += f"  {flops / ms / 1e9:8.1f} TFLOP/s"
    print(f"{tag:34s} {ms:8.3f} ms{extra}", flush=True)


x = torch.randn(T, h, device=dev).to(BF)
res = torch.randn(T, h, device=dev).to(BF)
w_up = (torch.randn(4 * h, h, device=dev) * 0.02).to(BF)
w_dn = (torch.randn(h, 4 * h, device=dev) * 0.02).to(BF)
w_qkv = (torch.randn(3 * h, h, device=dev) * 0.02).to(BF)
b_up, b_dn, b_qkv = torch.randn(4 * h, device=dev), torch.randn(h, device=dev), torch.randn(3 * h, device=dev)
h4 = torch.randn(T, 4 * h, device=dev).to(BF)
pre = torch.empty(T, 4 * h, dtype=BF, device=dev)
dw_up = torch.zeros(4 * h, h, device=dev)

run("gemm_fwd_bias(qkv)", lambda: K.gemm(x, w_qkv, bias=b_qkv), flops=2 * T * h * 3 * h)
run("gemm_fwd_bias_gelu_aux(mlp_up)", lambda: K.gemm(x, w_up, bias=b_up, gelu=True, aux_out=pre), flops=2 * T * h * 4 * h)
run("gemm_fwd_bias_residual(mlp_down)", lambda: K.gemm(h4, w_dn, bias=b_dn, residual=res), flops=2 * T * h * 4 * h)
run("gemm_dgrad_dgelu(mlp_down)", lambda: K.gemm(x, w_dn, b_mn=True, dgelu_in=pre), flops=2 * T * h * 4 * h)
run("gemm_dgrad_residual(mlp_up)", lambda: K.gemm(h4, w_up, b_mn=True, residual=res), flops=2 * T * h * 4 * h)
run("gemm_wgrad_splitk(mlp_up)", lambda: K.gemm(h4, x, a_mn=True, b_mn=True, out=dw_up, accumulate=True), flops=2 * T * h * 4 * h)
del h4, pre, dw_up, w_up, w_dn
Th = 8192  # LM head on a quarter of the tokens (the full [T, V] logits are 3.3 GB; same tile shapes)
w_out = (torch.randn(V, h, device=dev) * 0.02).to(BF)
run("gemm_lm_head(T/4)", lambda: K.gemm(x[:Th], w_out), flops=2 * Th * h * V)
logits = K.gemm(x[:Th], w_out)
labels = torch.randint(0, V, (Th,), device=dev)
run("cross_entropy_fwd_bwd(T/4)", lambda: K.cross_entropy_(logits, labels, V=V, write_grad=True), nbytes=2 * Th * V * 2)
del logits, w_out

# ---- attention, head_dim 256 (Pythia-1b)
qkv = (torch.randn(B, S, nh, 3, hd, device=dev) * 0.5).to(BF)
q, k, v = qkv[:, :, :, 0], qkv[:, :, :, 1], qkv[:, :, :, 2]
dqkv = torch.empty_like(qkv)
ang = torch.rand(S, 32, device=dev) * 6.2831853
cos, sin = ang.cos(), ang.sin()  # a real rotation: repeated in-place application must not grow q / k
run("rope_qk_inplace", lambda: K.rope_qk_inplace(qkv.view(T, -1), cos, sin, B, S, nh, hd, 64), nbytes=T * nh * 2 * 64 * 2 * 2)
fa = 4 * B * nh * S * S * hd / 2
run("attention_fwd_hd256", lambda: K.attention_fwd(q, k, v, causal=True), flops=fa)
o, lse = K.attention_fwd(q, k, v, causal=True)
do = torch.randn_like(o)
run("attention_bwd_hd256(delta+score+2 gemm)", lambda: K.attention_bwd(q, k, v, o, lse, do, dqkv[:, :, :, 0], dqkv[:, :, :, 1], dqkv[:, :, :, 2], causal=True), flops=2.5 * fa)
del qkv, dqkv, o, lse, do
K._score_cache.clear()

# ---- attention, head_dim 64 (Pythia-410m shape: 16 heads; RoBERTa uses the same kernels without the causal mask)
qkv = (torch.randn(B, S, 16, 3, 64, device=dev) * 0.5).to(BF)
q, k, v = qkv[:, :, :, 0], qkv[:, :, :, 1], qkv[:, :, :, 2]
dqkv = torch.empty_like(qkv)
fa64 = 4 * B * 16 * S * S * 64 / 2
run("attention_fwd_hd64", lambda: K.attention_fwd(q, k, v, causal=True), flops=fa64)
o, lse = K.attention_fwd(q, k, v, causal=True)
do = torch.randn_like(o)
run("attention_bwd_hd64(delta+dq+dkv)", lambda: K.attention_bwd(q, k, v, o, lse, do, dqkv[:, :, :, 0], dqkv[:, :, :, 1], dqkv[:, :, :, 2], causal=True), flops=2.5 * fa64)
del qkv, dqkv, o, lse, do

# ---- memory-bound kernels
g1, b1, g2, b2 = (torch.randn(h, device=dev) for _ in range(4))
run("layernorm_fwd_dual", lambda: K.layernorm_fwd(x, g1, b1, 1e-5, g2, b2), nbytes=T * h * 2 * 3)
y1, y2, mean, rstd = K.layernorm_fwd(x, g1, b1, 1e-5, g2, b2)
dy1, dy2 = torch.randn_like(x), torch.randn_like(x)
dg = [torch.zeros(h, device=dev) for _ in range(4)]
run("layernorm_bwd_dual_residual", lambda: K.layernorm_bwd(x, mean, rstd, g1, dy1, dg[0], dg[1], g2, dy2, dg[2], dg[3], dres=res), nbytes=T * h * 2 * 5)
cs = torch.zeros(3 * h, device=dev)
x3 = torch.randn(T, 3 * h, device=dev).to(BF)
run("colsum(bias grad, 3h)", lambda: K.colsum_(x3, cs), nbytes=T * 3 * h * 2)
del x3
table = (torch.randn(V, h, device=dev) * 0.02).to(BF)
ids = torch.randint(0, V, (T,), device=dev)
run("embedding_fwd", lambda: K.embedding_fwd(ids, table), nbytes=T * h * 2 * 2)
dtab = torch.zeros(V, h, device=dev)
run("embedding_bwd", lambda: K.embedding_bwd(ids, x, dtab), nbytes=T * h * (2 + 8))
run("dropout_residual", lambda: K.dropout(x, 0.1, 1234, residual=res), nbytes=T * h * 2 * 3)
del table, dtab, dy1, dy2, y1, y2

# ---- optimizer over a whole flat store (Pythia-1b: 1.01 B parameters; PROF_OPT_MODEL=pythia-410m keeps ncu replays short)
import os
cfg = pythia_config_dict(os.environ.get("PROF_OPT_MODEL", "pythia-1b"))
model = B200GPTNeoXForCausalLM(as_namespace(cfg)).to(dev)
opt = B200Adam(model.parameters(), lr=3e-4, betas=(0.9, 0.95))
n = model.flat.numel
model.flat.grad.normal_(0, 1e-3)
ss = torch.zeros((), device=dev)
run("sumsq(grad norm)", lambda: K.sumsq_(model.flat.grad, ss), nbytes=n * 4)
run(f"adam_step({n / 1e9:.2f}B params)", lambda: opt.step(), nbytes=n * 28)
run("cast_f32_to_bf16", lambda: K.cast_f32_to_bf16(model.flat.master, model.flat.shadow), nbytes=n * 6)
print("ok")
