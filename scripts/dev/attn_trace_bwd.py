"""Per-CTA phase clocks of the generic attention backward (dK/dV pass) for a RoBERTa-shaped problem (B200_ATTN_TRACE=3)."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from multimodal_llm_pretraining_b200 import _lib, kernels as K

dev = torch.device("cuda:0")
B, S, H, D = 64, 512, 16, 64
causal = False
qkv = torch.randn(B, S, 3, H, D, device=dev).to(torch.bfloat16)
q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
o, lse = K.attention_fwd(q, k, v, causal=causal)
d_o = torch.randn_like(o)
dqkv = torch.empty_like(qkv)
for _ in range(3):
    K.attention_bwd(q, k, v, o, lse, d_o, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2], causal=causal)
torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_ulonglong * 8192)()
lib.b200_debug_attn_trace.argtypes = [C.c_void_p, C.c_int]
assert lib.b200_debug_attn_trace(buf, 8192) == 0
e = [buf[8100 + i] for i in range(10)]
print("entry->setup", e[1] - e[0], "setup->s_full0", e[2] - e[1], "s_full0->s_full1", e[3] - e[2], "s_full1->loop_end", e[4] - e[3],
      "drain", e[5] - e[4], "epilogue", e[6] - e[5], "exit", e[7] - e[6], "total", e[7] - e[0], "ns", e[9] - e[8])
