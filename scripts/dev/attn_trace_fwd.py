"""Per-key-block event clocks of CTA (0,0,0) (the longest causal query tile) of the head_dim-256 forward (B200_ATTN_TRACE=1).
usage: B200_ATTN_TRACE=1 python scripts/dev/attn_trace_fwd.py [B] [H]"""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from multimodal_llm_pretraining_b200 import _lib, kernels as K

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1
S, D = 2048, 256
qkv = (torch.randn(B, S, H, 3, D, device=dev) * 0.5).to(torch.bfloat16)
q, k, v = qkv[:, :, :, 0], qkv[:, :, :, 1], qkv[:, :, :, 2]
for _ in range(3):
    K.attention_fwd(q, k, v, causal=True)
torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_ulonglong * 8192)()
lib.b200_debug_attn_trace.argtypes = [C.c_void_p, C.c_int]
assert lib.b200_debug_attn_trace(buf, 8192) == 0
t0 = buf[8000]
print(f"B {B} H {H}: CTA start 0, setup done {buf[8001] - t0}, Q in TMEM {buf[8002] - t0}, loop end {buf[8003] - t0}, epilogue end {buf[8004] - t0}, exit {buf[8005] - t0}; wall {buf[8007] - buf[8006]} ns")
prev = None
for j in range(32):
    e = {i: buf[4096 + 16 * j + i] - t0 for i in (0, 1, 8, 9, 10, 11, 12, 13)}
    per = "" if prev is None else f" period {e[13] - prev:5d}"
    prev = e[13]
    print(f"j={j:2d} MMA S_issue={e[0]:7d} PV_issue={e[1]:7d} | SMX top={e[8]:7d} s_full+{e[9] - e[8]:5d} ld+max+{e[10] - e[9]:5d} o_done_wait+{e[11] - e[10]:5d} exp+sts+{e[12] - e[11]:5d} fence+arrive+{e[13] - e[12]:4d}{per}")
