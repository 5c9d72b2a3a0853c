"""Dumps the per-tile event clocks of CTA (0,0,0) of the head_dim-256 score pass (run with B200_ATTN_TRACE=1)."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from multimodal_llm_pretraining_b200 import _lib, kernels as K

dev = torch.device("cuda:0")
B, S, H, D = 16, 2048, 8, 256
qkv = torch.randn(B, S, H, 3, D, device=dev).to(torch.bfloat16)
q, k, v = qkv[:, :, :, 0], qkv[:, :, :, 1], qkv[:, :, :, 2]
o, lse = K.attention_fwd(q, k, v, causal=True)
d_o = torch.randn_like(o)
dqkv = torch.empty_like(qkv)
for _ in range(3):
    K.attention_bwd(q, k, v, o, lse, d_o, dqkv[:, :, :, 0], dqkv[:, :, :, 1], dqkv[:, :, :, 2], causal=True)
torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_ulonglong * 8192)()
lib.b200_debug_attn_trace.argtypes = [C.c_void_p, C.c_int]
assert lib.b200_debug_attn_trace(buf, 8192) == 0
n_tiles = 32
t0 = buf[8]
names_m = ["s_free", "scores_issued", "a_ready", "dq_issued"]
names_c = ["top", "s_full", "ld_done", "math_done", "dq_done", "st_read", "sts_fence", "bar2"]
for t in range(n_tiles):
    m = [buf[16 * t + i] - t0 if buf[16 * t + i] else -1 for i in range(4)]
    c = [buf[16 * t + 8 + i] - t0 for i in range(8)]
    print(f"t={t:2d} MMA " + " ".join(f"{n}={x:7d}" for n, x in zip(names_m, m)) + " | CMP " + " ".join(f"{n}={x:7d}" for n, x in zip(names_c, c)))
