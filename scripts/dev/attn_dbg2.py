import sys, os
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from multimodal_llm_pretraining_b200 import kernels as K
dev = torch.device("cuda:0")
B, S, H, D = int(os.environ.get("B", "8")), 512, 16, 64
keep = []
for layer in range(24):
    qkv = torch.randn(B, S, 3, H, D, device=dev).to(torch.bfloat16)
    keep.append(qkv)
    keep.append(torch.empty(B * S, 4096, device=dev, dtype=torch.bfloat16))
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    seed = (1 * 1_000_003 + 4 * layer + 3) & 0xFFFFFFFFFFFF
    try:
        o, lse = K.attention_fwd(q, k, v, causal=False, dropout_p=0.1, dropout_seed=seed)
        torch.cuda.synchronize()
        keep.append(o)
    except Exception as e:
        print("FAIL layer", layer, "seed", seed, hex(qkv.data_ptr()), str(e)[:120])
        sys.exit(1)
    print("ok layer", layer, hex(qkv.data_ptr()))
print("all ok")
