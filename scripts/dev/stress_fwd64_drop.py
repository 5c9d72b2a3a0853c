"""Stand-alone stress of the head_dim-64 forward with dropout on the packed q|k|v layout RoBERTa uses."""
import os, sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from multimodal_llm_pretraining_b200 import kernels as K
dev = torch.device("cuda:0")
N = int(os.environ.get("N", "3000"))
B, S, H, D = 8, 512, 16, 64
torch.manual_seed(0)
qkv = (torch.randn(B, S, 3, H, D, device=dev) * 0.5).to(torch.bfloat16)
q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
i = -1
try:
    for i in range(N):
        K.attention_fwd(q, k, v, causal=False, dropout_p=0.1, dropout_seed=2000000 + i)
        if i % 100 == 99:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    print(f"stress_fwd64_drop OK: {N} launches")
except Exception as e:
    print(f"stress_fwd64_drop FAIL at launch ~{i}: {str(e)[:160]}")
