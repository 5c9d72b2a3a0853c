"""CPU tests of the reference-surface mirror: model registry, hyper-parameters, TrainingClass -> HF args dict (checked
against the one golden artefact the reference publishes: README.md:77-124), data sharding (bit exact), FLOP metric."""
import json
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from multimodal_llm_pretraining_b200.benchmarking.data import DummyTextModelingDataset, ShardedBatchIterator, shard_rows  # noqa: E402
from multimodal_llm_pretraining_b200.config import TrainingConfig  # noqa: E402
from multimodal_llm_pretraining_b200.models import get_model_class  # noqa: E402
from multimodal_llm_pretraining_b200.models import configs as C  # noqa: E402


def test_param_counts_match_published():
    for name, n in C.PYTHIA_PARAM_COUNTS.items():
        assert C.neox_param_count(C.pythia_config_dict(name)) == n, name


def test_flat_module_param_count_and_hf_keys():
    from multimodal_llm_pretraining_b200.modeling_gpt_neox import B200GPTNeoXForCausalLM

    cfg = C.pythia_config_dict("pythia-70m")
    m = B200GPTNeoXForCausalLM(C.as_namespace(cfg))
    assert sum(p.numel() for p in m.parameters()) == C.PYTHIA_PARAM_COUNTS["pythia-70m"]
    tr = pytest.importorskip("transformers")
    with torch.device("meta"):
        hf = tr.GPTNeoXForCausalLM(tr.GPTNeoXConfig(**cfg))
    hf_keys = sorted(k for k in hf.state_dict() if "inv_freq" not in k)
    assert sorted(m.state_dict()) == hf_keys
    for k, v in hf.state_dict().items():
        if "inv_freq" not in k:
            assert m.state_dict()[k].shape == v.shape, k
    # init statistics follow HF (N(0, 0.02), LN (1,0), bias 0)
    w = m.state_dict()["gpt_neox.layers.0.mlp.dense_h_to_4h.weight"]
    assert abs(w.std().item() - 0.02) < 2e-3 and abs(w.mean().item()) < 1e-3
    assert torch.all(m.state_dict()["gpt_neox.layers.0.input_layernorm.weight"] == 1)


def test_roberta_module_param_count_keys_and_flops():
    from multimodal_llm_pretraining_b200.modeling_roberta import B200RobertaForMaskedLM

    cfg = C.roberta_large_config_dict()
    small = dict(cfg, num_hidden_layers=2)
    m = B200RobertaForMaskedLM(C.as_namespace(small))
    tr = pytest.importorskip("transformers")
    with torch.device("meta"):
        hf = tr.RobertaForMaskedLM(tr.RobertaConfig(**small))
    assert sorted(m.state_dict()) == sorted(hf.state_dict())
    for k, v in hf.state_dict().items():
        assert m.state_dict()[k].shape == v.shape, k
    assert sum(p.numel() for p in m.parameters()) == sum(p.numel() for p in hf.parameters())
    # full roberta-large count (published) without building it: closed form from the shapes
    from multimodal_llm_pretraining_b200.modeling_roberta import roberta_param_shapes
    import math

    assert sum(math.prod(e[1]) for e in roberta_param_shapes(C.as_namespace(cfg))) == C.ROBERTA_LARGE_PARAM_COUNT
    # q/k/v weights and biases are contiguous in the flat store (one fused [3h, h] GEMM operand)
    f, h = m.flat, small["hidden_size"]
    p0 = "roberta.encoder.layer.0.attention.self"
    assert f.offsets[f"{p0}.key.weight"] - f.offsets[f"{p0}.query.weight"] == h * h
    assert f.offsets[f"{p0}.value.bias"] - f.offsets[f"{p0}.query.bias"] == 2 * h
    # W_lin and FLOPs per token of SURVEY.md section 8d
    assert C.roberta_linear_weight_count(cfg) == 303_038_464 + 51_471_360
    assert abs(C.roberta_train_flops_per_sequence(cfg, 512) / 512 / 1e9 - 2.278) < 2e-3
    assert abs(C.roberta_train_flops_per_sequence(cfg, 128) / 128 / 1e9 - 2.165) < 2e-3


def test_readme_training_arguments_golden():
    """scripts/to_training_arguments.py output for pythia-1b / free-lunch / zero_1 / mbs 16 / ga 16 (reference README.md:60-124)."""
    gold = json.load(open(ROOT / "tests" / "golden" / "readme_training_arguments.json"))
    cfg = TrainingConfig(num_nodes=1, gpus_per_node=4, gpu_type="a100", model="pythia-1b", free_lunch=True, sharding="zero_1")
    got = cfg.training_class(micro_batch_size=16, gradient_accumulation_steps=16)._to_huggingface_args_dict()
    assert json.loads(json.dumps(got)) == gold


def test_hyperparameters_follow_reference():
    p = get_model_class("pythia-1b")
    assert (p.batch_size, p.training_steps, p.mixed_precision, p.max_grad_norm) == (1024, 143000, "bf16", 1.0)
    assert p.optimizer_kwargs == {"lr": 3e-4, "betas": (0.9, 0.95), "eps": 1e-8, "weight_decay": 0.01}
    assert p.scheduler_type.value == "cosine_with_min_lr" and p.scheduler_kwargs == {"num_warmup_steps": 1430, "min_lr_rate": 0.1}
    assert (p.vocab_size, p.sequence_length) == (50304, 2049)
    assert get_model_class("pythia-160m").mixed_precision == "fp16" and get_model_class("pythia-160m").optimizer_kwargs["lr"] == 6e-4
    r = get_model_class("roberta")
    assert (r.batch_size, r.training_steps, r.max_grad_norm, r.vocab_size, r.sequence_length) == (8192, 500000, 0.0, 50265, 512)
    with pytest.raises(NotImplementedError):
        get_model_class("mamba")
    for sharding in ("zero_3", "fsdp_full_shard", "fsdp_hybrid_shard"):  # parameter sharding: out of scope
        tc = TrainingConfig(1, 8, "b200", "pythia-1b", sharding=sharding).training_class()
        assert tc.is_valid() and not tc.runs_on_b200_engine()
    for sharding in ("", "zero_1", "zero_2", "fsdp_shard_grad_op"):  # DDP, ZeRO-1, gradient sharding: built
        tc = TrainingConfig(1, 8, "b200", "pythia-1b", sharding=sharding).training_class()
        assert tc.is_valid() and tc.runs_on_b200_engine()
    assert not TrainingConfig(1, 8, "b200", "pythia-1b", sharding="zero_2", offloading=True).training_class().runs_on_b200_engine()


def test_flop_metric_closed_form():
    from multimodal_llm_pretraining_b200.benchmarking.flops import count_flops_per_example

    f = count_flops_per_example(get_model_class("pythia-1b"))
    assert abs(f - 12_817_874_812_992) / f < 1e-7  # FlopCounterMode value quoted in BASELINE.md
    assert abs(f / 2048 - 6.259e9) / 6.259e9 < 1e-3


def test_dummy_dataset_and_sharding_bit_exact():
    ds = DummyTextModelingDataset(vocab_size=50304, sequence_length=2049, num_samples=64, seed=7)
    assert ds.input_ids.dtype == torch.int64 and ds.input_ids.shape == (64, 2049)
    assert torch.equal(ds.input_ids, ds.labels) and ds.labels.data_ptr() != ds.input_ids.data_ptr()
    item = ds[3]
    assert set(item) == {"input_ids", "labels"} and int(ds.input_ids.max()) < 50304
    # every batch of 4 rows is consumed exactly once per epoch, round-robin over ranks
    W, mbs = 2, 4
    seen = []
    for t in range(64 // (W * mbs)):
        for r in range(W):
            seen.append(shard_rows(64, mbs, W, r, t))
    assert torch.equal(torch.cat(seen).sort().values, torch.arange(64))
    assert torch.equal(shard_rows(64, mbs, W, 1, 0), torch.arange(4, 8))
    assert torch.equal(shard_rows(64, mbs, W, 0, 1), torch.arange(8, 12))
    it0 = ShardedBatchIterator(ds, mbs, W, 0, shuffle_seed=0, pin=False)
    it1 = ShardedBatchIterator(ds, mbs, W, 1, shuffle_seed=0, pin=False)
    b0, b1 = next(it0), next(it1)
    perm = torch.randperm(64, generator=torch.Generator().manual_seed(0))
    assert torch.equal(b0["input_ids"], ds.input_ids[perm[0:4]]) and torch.equal(b1["input_ids"], ds.input_ids[perm[4:8]])
    assert torch.equal(b0["labels"], b0["input_ids"])


def test_analytic_training_days_with_b200_row():
    """experiments/training_time_analytic.py:14-53: days = training FLOPs / (GPUs x datasheet peak x 86400), b200 row added."""
    from multimodal_llm_pretraining_b200.benchmarking.flops import count_flops_per_example, estimate_training_days_from_flops
    from multimodal_llm_pretraining_b200.gpus import PEAK_TFLOPS, ampere_or_newer_gpu
    from multimodal_llm_pretraining_b200.models import get_model_class

    mc = get_model_class("pythia-1b")
    total = count_flops_per_example(mc) * mc.batch_size * mc.training_steps
    # FlopCounterMode's figure in SURVEY §8d (12 817 874 812 992) also counts the rotary inv_freq x position matmul, 2*32*2049 FLOPs
    assert abs(count_flops_per_example(mc) - 12_817_874_812_992) == 2 * 32 * 2049
    days = estimate_training_days_from_flops(1, 8, "b200", mc)
    assert days == pytest.approx(total / (8 * 2250e12 * 86400))
    assert estimate_training_days_from_flops(1, 8, "h100", mc) == pytest.approx(days * 2250 / 756)
    assert PEAK_TFLOPS["a100"] == {"bf16": 312.0, "tf32": 156.0} and ampere_or_newer_gpu("b200") and not ampere_or_newer_gpu("v100")
    with pytest.raises(NotImplementedError):
        estimate_training_days_from_flops(1, 8, "v100", mc)


def test_registry_matches_the_live_reference_registry():
    """When the reference checkout is present (this container, not the GPU box): every hyper-parameter property of every in-scope
    model class equals what the reference's own src/models registry returns (src/models/__init__.py:240-296, pythia.py, roberta.py)."""
    ref_root = Path("/root/reference")
    if not (ref_root / "src" / "models" / "__init__.py").exists():
        pytest.skip("reference checkout not present")
    from typing import get_args

    from multimodal_llm_pretraining_b200.models import PythiaT
    from multimodal_llm_pretraining_b200.optim import B200Adam, B200AdamW

    sys.path.insert(0, str(ref_root))
    try:
        from src.benchmarking.data import DummyTextModelingDataset as RefDataset
        from src.models import get_model_class as ref_get_model_class  # the reference's registry
    finally:
        sys.path.remove(str(ref_root))
    attrs = ["batch_size", "training_steps", "mixed_precision", "optimizer_kwargs", "scheduler_type", "scheduler_kwargs", "max_grad_norm",
             "hf_training_args", "fsdp_layers_to_wrap", "vocab_size", "sequence_length", "supports_activation_checkpointing", "supports_compilation"]
    for name in list(get_args(PythiaT)) + ["roberta"]:
        ref, ours = ref_get_model_class(name), get_model_class(name)
        for a in attrs:
            mine, theirs = getattr(ours, a), getattr(ref, a)
            assert str(getattr(mine, "value", mine)) == str(getattr(theirs, "value", theirs)), (name, a, mine, theirs)
        # the registry hands out the reference's own class; the trainer maps it onto the fused equivalent for a B200 module
        assert ours.optimizer is ref.optimizer, name
        from multimodal_llm_pretraining_b200.optim import fused_optimizer_class
        assert fused_optimizer_class(ours.optimizer) is {torch.optim.Adam: B200Adam, torch.optim.AdamW: B200AdamW}[ref.optimizer], name
    # the synthetic dataset: same constructor arguments, item keys, dtypes and shapes (src/benchmarking/data.py:8-21)
    ref_ds = RefDataset(vocab_size=50304, sequence_length=2049, num_samples=8)
    our_ds = get_model_class("pythia-70m").load_dummy_dataset(num_samples=8, seed=0)
    assert len(ref_ds) == len(our_ds) == 8
    assert {k: (v.dtype, v.shape) for k, v in ref_ds[0].items()} == {k: (v.dtype, v.shape) for k, v in our_ds[0].items()}


def test_print_optimal_config_table(tmp_path):
    """scripts/print_optimal_config.py:8-48 of the reference: rows of one (nodes, gpus, gpu type, model), sorted by training days,
    with grad_acc_steps = batch_size // (micro_batch_size * gpus_per_node); same columns in the same order."""
    import json
    import sys as _sys

    _sys.path.insert(0, str(ROOT / "scripts"))
    import print_optimal_config as P

    rows = [
        dict(num_nodes=1, gpus_per_node=8, gpu_type="b200", model="pythia-1b", free_lunch=True, activation_checkpointing=False, sharding="zero_1", offloading=False, micro_batch_size=16, training_days=2.41),
        dict(num_nodes=1, gpus_per_node=8, gpu_type="b200", model="pythia-1b", free_lunch=True, activation_checkpointing=True, sharding="", offloading=False, micro_batch_size=64, training_days=3.3),
        dict(num_nodes=1, gpus_per_node=8, gpu_type="b200", model="pythia-1b", free_lunch=True, activation_checkpointing=False, sharding="", offloading=False, micro_batch_size=16, training_days=2.38),
        dict(num_nodes=1, gpus_per_node=8, gpu_type="b200", model="pythia-1b", free_lunch=True, activation_checkpointing=False, sharding="", offloading=False, micro_batch_size=16, training_days=2.37),  # re-run replaces
        dict(num_nodes=1, gpus_per_node=8, gpu_type="b200", model="pythia-1b", free_lunch=True, activation_checkpointing=False, sharding="zero_3", offloading=False, micro_batch_size=0, training_days=None),
        dict(num_nodes=1, gpus_per_node=4, gpu_type="b200", model="pythia-1b", free_lunch=True, activation_checkpointing=False, sharding="", offloading=False, micro_batch_size=16, training_days=4.7),
    ]
    f = tmp_path / "r.jsonl"
    f.write_text("\n".join(json.dumps(r) for r in rows))
    t = P.optimal_config_table(P.load_results(f), 1, 8, "b200", "pythia-1b")
    assert list(t[0]) == ["num_nodes", "gpus_per_node", "gpu_type", "model", "free_lunch", "activation_checkpointing", "sharding", "offloading",
                          "micro_batch_size", "grad_acc_steps", "training_days"]
    assert [r["training_days"] for r in t] == [2.37, 2.41, 3.3]
    assert [r["grad_acc_steps"] for r in t] == [8, 8, 2]  # 1024 // (mbs * 8)
    assert "grad_acc_steps" in P.format_table(t) and P.format_table([]).startswith("(no benchmarked")
