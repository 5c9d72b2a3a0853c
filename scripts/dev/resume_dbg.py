"""Triage: is a checkpoint resume bit-exact on one GPU? Compares restored master / shadow / moments with the live ones and the
step taken from the restored state with the uninterrupted run."""
import sys, tempfile
sys.path.insert(0, "/root/repo")
import torch
from multimodal_llm_pretraining_b200.engine import TrainEngine
from multimodal_llm_pretraining_b200.modeling_gpt_neox import B200GPTNeoXForCausalLM
from multimodal_llm_pretraining_b200.models.configs import as_namespace, pythia_config_dict
from multimodal_llm_pretraining_b200.optim import B200Adam
dev = torch.device("cuda:0")
cfg = dict(pythia_config_dict("pythia-70m"), num_hidden_layers=2, vocab_size=1024)
def tiny(seed=0):
    m = B200GPTNeoXForCausalLM(as_namespace(cfg)); m.reset_parameters(torch.Generator().manual_seed(seed)); return m.to(dev).train()
def eng(m): return TrainEngine(m, B200Adam(m.parameters(), lr=1e-3, betas=(0.9, 0.95)), None, max_grad_norm=1.0, gradient_accumulation_steps=2, strategy="none")
data = torch.randint(0, 1024, (4, 2, 4, 129), generator=torch.Generator().manual_seed(11)).to(dev)
def run(e, steps):
    for s in steps:
        for mb in range(2): e.manual_training_step({"input_ids": data[s, mb], "labels": data[s, mb]})
        e.manual_optimization_step()
a = tiny(); ea = eng(a); run(ea, range(3))
b = tiny(); eb = eng(b); run(eb, range(2)); two = b.flat.master.clone(); m2, v2 = eb.optimizer._m.clone(), eb.optimizer._v.clone()
d = tempfile.mkdtemp(); eb.save_checkpoint(d)
c = tiny(99); ec = eng(c); ec.load_checkpoint(d)
print("master equal after load", torch.equal(c.flat.master, two), "shadow", torch.equal(c.flat.shadow, b.flat.shadow), "m", torch.equal(ec.optimizer._m, m2), "v", torch.equal(ec.optimizer._v, v2), "step", ec.optimizer._step)
# where do master buffers differ (padding?)
diff = (c.flat.master != two).nonzero().flatten()
print("n differing master elems", diff.numel(), diff[:5].tolist())
sdiff = (c.flat.shadow != b.flat.shadow).nonzero().flatten()
print("n differing shadow elems", sdiff.numel(), sdiff[:5].tolist())
run(eb, range(2, 3)); run(ec, range(2, 3))
def rel(x, y, base): return (((x - base) - (y - base)).norm() / (y - base).norm()).item()
print("continue vs straight", rel(b.flat.master, a.flat.master, two), "resumed vs straight", rel(c.flat.master, a.flat.master, two))
for n, p in c.named_parameters():
    pa = dict(a.named_parameters())[n]; p2 = a.flat.view(two, n)
    e = rel(p.data, pa.data, p2)
    if e > 1e-6: print("  ", n, e)
