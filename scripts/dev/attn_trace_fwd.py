"""Dumps per-tile event clocks of CTA (0,0,0) of the head_dim-256 forward kernel (run with B200_ATTN_TRACE=1)."""
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
for _ in range(3):
    o, lse = K.attention_fwd(q, k, v, causal=True)
torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_ulonglong * 8192)()
lib.b200_debug_attn_trace.argtypes = [C.c_void_p, C.c_int]
assert lib.b200_debug_attn_trace(buf, 8192) == 0
t0 = buf[4096 + 8]
for t in range(32):
    g = lambda i: (buf[4096 + 16 * t + i] - t0) if buf[4096 + 16 * t + i] else -1
    print(f"j={t:2d} MMA s_issue={g(0):7d} pv_issue={g(1):7d} | SM top={g(8):7d} s_full={g(9):7d} max_done={g(10):7d} pbuf_free={g(11):7d} exp_done={g(12):7d} p_ready={g(13):7d}")

e = [buf[8000 + i] for i in range(8)]
print("entry->setup", e[1] - e[0], "setup->q_ready", e[2] - e[1], "q_ready->loop_end", e[3] - e[2], "epilogue", e[4] - e[3], "exit", e[5] - e[4],
      "total cycles", e[5] - e[0], "total ns", e[7] - e[6])
