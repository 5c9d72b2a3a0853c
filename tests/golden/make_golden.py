"""Generates tests/golden/neox_tiny.pt by running the REAL reference call path on CPU in this container:
transformers.GPTNeoXForCausalLM (what src/models/pythia.py:15-22 builds) + torch.optim.Adam with the reference's
hyper-parameters (src/models/pythia.py:43-67) + clip_grad_norm_(1.0) (src/benchmarking/utils.py:66-70), fp32.

Run:  python tests/golden/make_golden.py      (needs transformers; not needed at test time)
"""
import sys
from pathlib import Path

import torch
from transformers import GPTNeoXConfig, GPTNeoXForCausalLM

OUT = Path(__file__).resolve().parent / "neox_tiny.pt"

CFG = dict(vocab_size=256, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=512,
           max_position_embeddings=128, rotary_pct=0.25, rotary_emb_base=10000, layer_norm_eps=1e-5,
           use_parallel_residual=True, hidden_act="gelu", attention_bias=True, hidden_dropout=0.0,
           attention_dropout=0.0, tie_word_embeddings=False, initializer_range=0.02)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(4)
    cfg = GPTNeoXConfig(**CFG, attn_implementation="eager")
    model = GPTNeoXForCausalLM(cfg).float()
    # make biases / LN affine non-trivial so the fixture exercises them
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith(".bias"):
                p.normal_(0, 0.02)
            elif "layernorm.weight" in n or "layer_norm.weight" in n:
                p.add_(torch.randn_like(p) * 0.05)
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items() if "inv_freq" not in k}
    g = torch.Generator().manual_seed(1)
    batches = [torch.randint(0, CFG["vocab_size"], (2, 65), generator=g) for _ in range(3)]

    # single fwd+bwd
    out = model(input_ids=batches[0], labels=batches[0])
    out.loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
    logits0 = out.logits.detach()[:, :4, :8].clone()
    loss0 = out.loss.item()
    model.zero_grad()

    # three optimizer steps, reference hyper-parameters (pythia-160m lr), warmup-free constant lr for the fixture
    opt = torch.optim.Adam(model.parameters(), lr=6e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.0)
    losses, norms = [], []
    for ids in batches:
        loss = model(input_ids=ids, labels=ids).loss
        loss.backward()
        norms.append(torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0).item())
        opt.step()
        model.zero_grad()
        losses.append(loss.item())
    sd3 = {k: v.detach().clone() for k, v in model.state_dict().items() if "inv_freq" not in k}
    keep = ["gpt_neox.layers.0.attention.query_key_value.weight", "gpt_neox.layers.1.mlp.dense_4h_to_h.bias",
            "gpt_neox.layers.1.input_layernorm.weight", "gpt_neox.final_layer_norm.bias", "embed_out.weight"]
    torch.save({
        "cfg": CFG, "state_dict": sd0, "batches": batches, "loss0": loss0, "logits0_slice": logits0,
        "grad_norms": {k: v.norm().item() for k, v in grads.items()},
        "grads": {k: grads[k] for k in keep},
        "losses": losses, "clip_norms": norms,
        "param_norms_after3": {k: v.norm().item() for k, v in sd3.items()},
        "params_after3": {k: sd3[k] for k in keep},
        "versions": {"torch": torch.__version__, "transformers": __import__("transformers").__version__},
    }, OUT)
    print("wrote", OUT, OUT.stat().st_size, "bytes; loss0", loss0, "losses", losses)


if __name__ == "__main__":
    sys.exit(main())
