"""Generates tests/golden/roberta_tiny.pt from the REAL reference call path on CPU in this container:
transformers.RobertaForMaskedLM (what src/models/roberta.py:15-18 builds; eager attention, dropout set to 0 for parity)
+ torch.optim.Adam with the reference's hyper-parameters (src/models/roberta.py:33-41), fp32, max_grad_norm 0 (no clipping,
src/models/roberta.py:52-54).   Run:  python tests/golden/make_golden_roberta.py
"""
from pathlib import Path

import torch
from transformers import RobertaConfig, RobertaForMaskedLM

OUT = Path(__file__).resolve().parent / "roberta_tiny.pt"
CFG = dict(vocab_size=301, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
           max_position_embeddings=66, type_vocab_size=1, pad_token_id=1, layer_norm_eps=1e-5, hidden_dropout_prob=0.0,
           attention_probs_dropout_prob=0.0, hidden_act="gelu", initializer_range=0.02)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(4)
    model = RobertaForMaskedLM(RobertaConfig(**CFG, attn_implementation="eager")).float().train()
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith(".bias") or n == "lm_head.bias":
                p.normal_(0, 0.02)
            elif "LayerNorm.weight" in n or "layer_norm.weight" in n:
                p.add_(torch.randn_like(p) * 0.05)
    tied = ("lm_head.decoder.weight", "lm_head.decoder.bias")  # aliases of the word embeddings / lm_head.bias
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items() if k not in tied}
    g = torch.Generator().manual_seed(1)
    batches = [torch.randint(0, CFG["vocab_size"], (2, 64), generator=g) for _ in range(3)]
    batches[0][0, 5] = 1  # pad tokens exercise the position-id rule
    batches[0][1, 0] = 1
    out = model(input_ids=batches[0], labels=batches[0])
    out.loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
    rec = dict(cfg=CFG, state_dict=sd0, batches=batches, loss0=out.loss.detach().clone(), logits0=out.logits.detach().clone(), grads0=grads)
    model.zero_grad()
    opt = torch.optim.Adam(model.parameters(), lr=4e-4, betas=(0.9, 0.98), weight_decay=0.0)
    losses = []
    for b in batches:
        loss = model(input_ids=b, labels=b).loss
        loss.backward()
        opt.step()
        model.zero_grad()
        losses.append(loss.detach().clone())
    rec["losses_3steps"] = torch.stack(losses)
    rec["state_dict_after3"] = {k: v.detach().clone() for k, v in model.state_dict().items() if k not in tied}
    torch.save(rec, OUT)
    print("wrote", OUT, {k: (v.shape if hasattr(v, "shape") else type(v)) for k, v in rec.items() if k in ("loss0", "logits0")})


if __name__ == "__main__":
    main()
