"""scripts/training.py on the B200: a real text-LM training run driven by a training-arguments JSON of the
shape scripts/to_training_arguments.py writes; interrupted + resumed == uninterrupted, bit for bit (data order, LR schedule,
optimizer moments, dropout step seed all restored): the logged losses of the steps after the resume are identical and the final
parameters agree to the run-to-run noise of the fp32 atomics (embedding backward, split-K wgrad partial sums land in arbitrary
order; Adam's first steps turn that rounding noise into +-lr steps on the few elements whose gradient is ~0)."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "scripts"))


def _args(**over):
    a = dict(max_steps=4, per_device_train_batch_size=2, gradient_accumulation_steps=2, lr_scheduler_type="cosine_with_min_lr",
             lr_scheduler_kwargs={"min_lr_rate": 0.1}, warmup_steps=2, gradient_checkpointing=False, bf16=True, fp16=False, tf32=True,
             fsdp="", fsdp_config=None, deepspeed=None, ddp_find_unused_parameters=False, torch_compile=True, max_grad_norm=1.0,
             logging_steps=1, save_steps=100, seed=5)
    a.update(over)
    return a


@pytest.mark.parametrize("model_type", ["pythia-70m", "roberta"])
def test_training_run_resumes_bit_exact(tmp_path, model_type, monkeypatch):
    import training as T
    from multimodal_llm_pretraining_b200.models import get_model_class
    from multimodal_llm_pretraining_b200.models.configs import as_namespace

    if model_type == "roberta":
        # roberta-large is 355 M parameters x 24 layers: shrink the depth for the test, keep widths / vocabulary / both dropouts
        import multimodal_llm_pretraining_b200.models.roberta as R
        from multimodal_llm_pretraining_b200.modeling_roberta import B200RobertaForMaskedLM

        def small(self, use_custom_kernels=True):
            return B200RobertaForMaskedLM(as_namespace(dict(self.config_dict(), num_hidden_layers=2)))

        monkeypatch.setattr(R.RobertaModelClass, "build_model", small)
    mc = get_model_class(model_type)
    S = mc.sequence_length
    rng = np.random.default_rng(0)
    (rng.integers(3, mc.vocab_size - 1, size=S * 24, dtype=np.int64).astype(np.uint16)).tofile(tmp_path / "train.bin")
    straight = T.train(str(tmp_path / "a"), model_type, _args(), tmp_path, "train")
    assert straight["global_step"] == 4 and len(straight["log_history"]) == 4
    losses = [r["loss"] for r in straight["log_history"]]
    assert all(np.isfinite(losses)) and abs(losses[0] - np.log(mc.vocab_size)) < 0.6
    assert (tmp_path / "a" / "checkpoint-4" / "pytorch_model.bin").exists()
    final = straight["trainer"].model.flat.master.clone()
    torch.manual_seed(5)
    init = T.get_model(model_type).flat.master.to(final.device)
    lrs = [r["learning_rate"] for r in straight["log_history"]]
    assert lrs[0] < lrs[1], "warm-up"
    del straight
    T.train(str(tmp_path / "b"), model_type, _args(_stop_after_steps=2), tmp_path, "train")
    assert (tmp_path / "b" / "checkpoint-2" / "trainer_state.json").exists()
    resumed = T.train(str(tmp_path / "b"), model_type, _args(resume_from_checkpoint=True), tmp_path, "train")
    assert resumed["global_step"] == 4
    assert [r["step"] for r in resumed["log_history"]] == [1, 2, 3, 4], "the log history continues"
    for a, b in zip(resumed["log_history"], [dict(loss=x) for x in losses]):
        assert abs(a["loss"] - b["loss"]) <= 1e-5 * abs(b["loss"]), (resumed["log_history"], losses)  # same data, same masks, same state
    got = resumed["trainer"].model.flat.master
    err = ((got - final).norm() / (final - init).norm()).item()
    print(f"{model_type}: interrupted + resumed vs uninterrupted, update rel diff {err:.3e} (tolerance 2e-2 = atomic-order noise through Adam)")
    assert err <= 2e-2, err


def test_fp16_training_run_uses_loss_scaling(tmp_path):
    import training as T
    from multimodal_llm_pretraining_b200.models import get_model_class

    mc = get_model_class("pythia-70m")
    rng = np.random.default_rng(1)
    (rng.integers(0, mc.vocab_size, size=mc.sequence_length * 16, dtype=np.int64).astype(np.uint16)).tofile(tmp_path / "train.bin")
    res = T.train(str(tmp_path / "o"), "pythia-70m", _args(bf16=False, fp16=True, max_steps=3), tmp_path, "train")
    tr = res["trainer"]
    assert tr.model.flat.shadow.dtype == torch.float16 and tr.engine.loss_scaler is not None
    assert all(r["loss_scale"] is not None and np.isfinite(r["loss"]) for r in res["log_history"])
    assert any(r["optimizer_step_taken"] for r in res["log_history"])
