"""Run-to-run reproducibility of one forward+backward on identical inputs: per-parameter relative difference of the gradients
between repeated runs. Expected: exact for everything except buffers reached by fp32 atomics (wgrad split-K red.add, embedding
backward), which differ at the 1e-7 level. Anything larger means a kernel reads memory it did not write."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from multimodal_llm_pretraining_b200.modeling_gpt_neox import B200GPTNeoXForCausalLM
from multimodal_llm_pretraining_b200.models.configs import as_namespace, pythia_config_dict

dev = torch.device("cuda:0")
cases = {
    "hd64 (70m, 3 layers)": (dict(pythia_config_dict("pythia-70m"), num_hidden_layers=3, vocab_size=2048), (4, 257)),
    "hd128 (h256, 2 heads)": (dict(pythia_config_dict("pythia-70m"), num_hidden_layers=2, vocab_size=2048, hidden_size=256, num_attention_heads=2, intermediate_size=1024), (4, 257)),
    "hd256 (h512, 2 heads, S 512)": (dict(pythia_config_dict("pythia-70m"), num_hidden_layers=2, vocab_size=2048, hidden_size=512, num_attention_heads=2, intermediate_size=2048), (4, 513)),
    "hd80 (h320, 4 heads)": (dict(pythia_config_dict("pythia-70m"), num_hidden_layers=2, vocab_size=2048, hidden_size=320, num_attention_heads=4, intermediate_size=1280), (4, 257)),
}
for name, (cfg, shape) in cases.items():
    m = B200GPTNeoXForCausalLM(as_namespace(cfg))
    m.reset_parameters(torch.Generator().manual_seed(0))
    m = m.to(dev).train()
    ids = torch.randint(0, 2048, shape, generator=torch.Generator().manual_seed(1)).to(dev)
    runs = []
    for r in range(3):
        # churn the allocator so that scratch buffers land on different (dirty) memory between runs
        junk = [torch.full((1 << 20,), float("nan"), device=dev) for _ in range(r * 3)]
        del junk
        m.zero_grad()
        loss = m(input_ids=ids, labels=ids)["loss"]
        loss.backward()
        torch.cuda.synchronize()
        runs.append((loss.item(), {n: p.grad.clone() for n, p in m.named_parameters()}))
    worst = (0.0, "")
    for n in runs[0][1]:
        for r in (1, 2):
            a, b = runs[0][1][n], runs[r][1][n]
            d = ((a - b).norm() / (a.norm() + 1e-30)).item()
            if not (d <= worst[0]):
                worst = (d, n)
    print(f"{name:30s} loss {[x[0] for x in runs]}  worst grad rel diff {worst[0]:.3e} ({worst[1]})", flush=True)
