"""Step engine and harness on the B200 (single GPU): checkpoint / resume, the reference's benchmarking flow
(find_max_mbs_pow2 -> benchmark_acc_optim_times -> estimate_step_time -> training days; src/benchmarking/*.py,
experiments/training_time_empirical.py:43-138), and the regression test of the head_dim-256 forward deadlock."""
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from multimodal_llm_pretraining_b200 import kernels as K  # noqa: E402
from multimodal_llm_pretraining_b200.engine import TrainEngine  # noqa: E402
from multimodal_llm_pretraining_b200.modeling_gpt_neox import B200GPTNeoXForCausalLM  # noqa: E402
from multimodal_llm_pretraining_b200.models.configs import as_namespace, pythia_config_dict  # noqa: E402
from multimodal_llm_pretraining_b200.optim import B200Adam, get_scheduler  # noqa: E402


def _tiny(dev, seed=0):
    cfg = dict(pythia_config_dict("pythia-70m"), num_hidden_layers=2, vocab_size=1024)
    m = B200GPTNeoXForCausalLM(as_namespace(cfg))
    m.reset_parameters(torch.Generator().manual_seed(seed))
    return m.to(dev).train()


def _engine(model, ga=2):
    opt = B200Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.95))
    sched = get_scheduler("cosine_with_min_lr", opt, 3, 100, {"min_lr_rate": 0.1})
    return TrainEngine(model, opt, sched, max_grad_norm=1.0, gradient_accumulation_steps=ga, strategy="none")


def test_checkpoint_resume_continues_the_same_trajectory(dev, tmp_path):
    """4 steps in one go == 2 steps, save_checkpoint, fresh objects, load_checkpoint, 2 more steps (parameters, Adam moments,
    step count and LR schedule all resume). Tolerance: the wgrad reductions accumulate through red.add in arbitrary order,
    so two runs of the same step differ at the 1e-6 level of the fp32 update."""
    g = torch.Generator().manual_seed(11)
    data = torch.randint(0, 1024, (4, 2, 4, 129), generator=g).to(dev)  # [step, micro, mbs, S]

    def run(eng, steps):
        for s in steps:
            for mb in range(2):
                eng.manual_training_step({"input_ids": data[s, mb], "labels": data[s, mb]})
            eng.manual_optimization_step()

    ref_model = _tiny(dev)
    ref = _engine(ref_model)
    run(ref, range(4))

    m1 = _tiny(dev)
    e1 = _engine(m1)
    run(e1, range(2))
    e1.save_checkpoint(tmp_path / "ckpt")
    assert {p.name for p in (tmp_path / "ckpt").iterdir()} == {"pytorch_model.bin", "optimizer.pt", "scheduler.pt", "trainer_state.json"}
    sd = torch.load(tmp_path / "ckpt" / "pytorch_model.bin")
    assert "gpt_neox.layers.0.attention.query_key_value.weight" in sd and "embed_out.weight" in sd  # HF key names

    m2 = _tiny(dev, seed=123)  # different init: everything must come from the checkpoint
    e2 = _engine(m2)
    e2.load_checkpoint(tmp_path / "ckpt")
    assert e2.micro == 4 and e2.optimizer._step == 2
    assert e2.scheduler.last_epoch == 2 and e2.optimizer.param_groups[0]["lr"] == e1.optimizer.param_groups[0]["lr"]
    assert torch.equal(m2.flat.master, m1.flat.master)
    assert torch.equal(m2.flat.shadow, m1.flat.master.to(torch.bfloat16))
    run(e2, range(2, 4))
    upd = m2.flat.master - m1.flat.master
    upd_ref = ref_model.flat.master - m1.flat.master
    err = ((upd - upd_ref).norm() / upd_ref.norm()).item()
    assert err <= 1e-2, err  # a lost moment buffer or step count shows up as O(1)
    assert e2.optimizer.param_groups[0]["lr"] == ref.optimizer.param_groups[0]["lr"]


def test_reference_benchmark_flow_single_gpu(dev):
    """TrainingConfig -> TrainingClass -> ManualTrainer, then the reference's three benchmark steps on a 2-layer model."""
    from multimodal_llm_pretraining_b200.benchmarking.max_batch_size import find_max_mbs_pow2
    from multimodal_llm_pretraining_b200.benchmarking.step_time import benchmark_acc_optim_times, compute_training_days, estimate_step_time
    from multimodal_llm_pretraining_b200.config import TrainingConfig

    config = TrainingConfig(1, 1, "b200", "pythia-70m", free_lunch=True)
    tc = config.training_class(num_training_steps=1, micro_batch_size=1, gradient_accumulation_steps=1, bf16=True, fp16=False)
    assert tc.is_valid() and tc.runs_on_b200_engine()
    mc = config.model_class()
    model = _tiny(dev)
    ds = mc.load_dummy_dataset(num_samples=64, seed=0)
    ds.input_ids %= 1024
    ds.labels %= 1024
    trainer = tc.build_trainer(model, ds)
    assert find_max_mbs_pow2(trainer, limit=4) == 4
    acc, opt = benchmark_acc_optim_times(trainer, micro_batch_size=2, training_steps=2, accumulations=2, warmup=True)
    assert 0 < acc < 1.0 and 0 < opt < 1.0
    step = estimate_step_time(trainer, micro_batch_size=2, target_micro_batch_size=8, num_benchmarking_steps=2)
    assert step > 3 * acc * 0.5  # 4 accumulations + the optimizer step
    assert compute_training_days(143_000, 1.0) == pytest.approx(143_000 / 86_400)


def test_oom_surfaces_as_torch_out_of_memory(dev):
    """find_max_mbs_pow2 relies on allocation failure raising torch.cuda.OutOfMemoryError (src/benchmarking/max_batch_size.py:18):
    every buffer of the step is a torch allocation, so it does."""
    m = _tiny(dev)
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    filler = torch.empty(free - (192 << 20), dtype=torch.uint8, device=dev)  # leave ~192 MB
    ids = torch.randint(0, 1024, (64, 1025), device=dev)  # 65k tokens: the qkv activations of one layer alone need 200 MB
    with pytest.raises(torch.cuda.OutOfMemoryError):
        m(input_ids=ids, labels=ids)["loss"]
    del filler
    m.zero_grad()
    torch.cuda.empty_cache()
    small = torch.randint(0, 1024, (2, 65), device=dev)
    assert torch.isfinite(m(input_ids=small, labels=small)["loss"])  # the module is still usable afterwards


def test_attention_fwd_hd256_survives_late_tiles(dev):
    """Regression: with ONE p_ready barrier for both P buffers, a V tile landing later than one softmax iteration left the MMA
    issuer two barrier phases behind and the CTA deadlocked (about once per 10^6 CTAs in a training step; within 200 launches
    when a second stream saturates HBM). 1500 launches with that second stream must complete."""
    B, S, H, D = 16, 2048, 8, 256
    qkv = (torch.randn(B, S, H, 3, D, device=dev, generator=torch.Generator(device=dev).manual_seed(0)) * 0.5).to(torch.bfloat16)
    q, k, v = qkv[:, :, :, 0], qkv[:, :, :, 1], qkv[:, :, :, 2]
    big = torch.empty(1 << 28, dtype=torch.float32, device=dev)
    side = torch.cuda.Stream()
    o_ref, lse_ref = K.attention_fwd(q, k, v, causal=True)
    torch.cuda.synchronize()
    for i in range(1500):
        if i % 4 == 0:
            with torch.cuda.stream(side):
                big.add_(1.0)
        o, lse = K.attention_fwd(q, k, v, causal=True)
        if i % 250 == 249:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    assert torch.equal(o, o_ref) and torch.equal(lse, lse_ref)
