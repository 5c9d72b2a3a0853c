import os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from multimodal_llm_pretraining_b200.models import get_model_class
from multimodal_llm_pretraining_b200.models.configs import as_namespace
from multimodal_llm_pretraining_b200.modeling_roberta import B200RobertaForMaskedLM
dev = torch.device("cuda:0")
mc = get_model_class("roberta")
cfg = dict(mc.config_dict())
cfg["num_hidden_layers"] = int(os.environ.get("L", "2"))
if os.environ.get("PATTN") is not None:
    cfg["attention_probs_dropout_prob"] = float(os.environ["PATTN"])
if os.environ.get("PHID") is not None:
    cfg["hidden_dropout_prob"] = float(os.environ["PHID"])
torch.manual_seed(0)
m = B200RobertaForMaskedLM(as_namespace(cfg)).to(dev).train()
B = int(os.environ.get("B", "8"))
ids = torch.randint(0, mc.vocab_size, (B, 512), generator=torch.Generator().manual_seed(1)).to(dev)
from multimodal_llm_pretraining_b200 import kernels as K
_orig = K.attention_fwd
calls = [0]
def _wrapped(q, k, v, causal, scale=None, dropout_p=0.0, dropout_seed=0):
    calls[0] += 1
    try:
        r = _orig(q, k, v, causal, scale, dropout_p, dropout_seed)
        torch.cuda.synchronize()
        return r
    except Exception:
        print("attention_fwd call", calls[0], "seed", dropout_seed, "q ptr", hex(q.data_ptr()), "strides", q.stride(), "p", dropout_p)
        raise
K.attention_fwd = _wrapped
try:
    for it in range(2):
        loss = m(input_ids=ids, labels=ids)["loss"]
        loss.backward()
        torch.cuda.synchronize()
        print("iter", it, "loss", loss.item())
    print("OK")
except Exception as e:
    print("FAIL", str(e)[:200])
