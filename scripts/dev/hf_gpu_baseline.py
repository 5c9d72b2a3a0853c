"""The nearest runnable stand-in for the reference's own GPU path on this box (SURVEY §8d "GPU comparison"): stock
transformers GPTNeoXForCausalLM (sdpa attention, bf16 autocast — what src/models/pythia.py:15-22 builds with
use_custom_kernels=True and src/train.py:94-124 configures) + torch.optim.Adam(fused=True) + clip_grad_norm_, same
synthetic batches and step definition as bench.py (grad-acc micro-batches of mbs x 2049 tokens, device-timed).
accelerate / DeepSpeed / HF Trainer are not in the image, so this is a plain loop over the same calls.
usage: hf_gpu_baseline.py [model] [mbs] [grad_acc] [steps]"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from multimodal_llm_pretraining_b200.models.configs import neox_train_flops_per_sequence, pythia_config_dict

model_name = sys.argv[1] if len(sys.argv) > 1 else "pythia-1b"
mbs = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ga = int(sys.argv[3]) if len(sys.argv) > 3 else 4
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
from transformers import GPTNeoXConfig, GPTNeoXForCausalLM

dev = torch.device("cuda:0")
cfg = pythia_config_dict(model_name)
torch.manual_seed(0)
model = GPTNeoXForCausalLM(GPTNeoXConfig(**cfg, attn_implementation="sdpa")).to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=3e-4, betas=(0.9, 0.95), eps=1e-8, fused=True)
S = 2049
batches = [torch.randint(0, cfg["vocab_size"], (mbs, S), device=dev) for _ in range(ga)]


def step():
    for ids in batches:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = model(input_ids=ids, labels=ids).loss
        (loss / ga).backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    opt.step()
    opt.zero_grad(set_to_none=True)
    return loss


step()  # warm-up
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
tok = ga * mbs * (S - 1)
f_tok = neox_train_flops_per_sequence(cfg, S) / (S - 1)
print(json.dumps({"impl": "hf-transformers-gpu (sdpa, bf16 autocast, fused Adam)", "model": model_name, "micro_batch": mbs, "grad_acc": ga,
                  "ms_per_step": ms, "tokens_per_s": tok / ms * 1e3, "mfu_vs_2250": tok / ms * 1e3 * f_tok / 2250e12,
                  "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9, "last_loss": float(loss)}))
