"""CPU tests of the text-LM data path behind scripts/training.py: token-file dataset, causal / masked-LM
collators (HF DataCollatorForLanguageModeling semantics), the resumable per-rank sampler, the argument plumbing."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "scripts"))

from multimodal_llm_pretraining_b200.text_data import CausalLMCollator, EpochSampler, MaskedLMCollator, TokenFileDataset  # noqa: E402


def test_token_file_dataset_bin_npy_and_split_dir(tmp_path):
    toks = np.arange(1000, dtype=np.uint16) % 997
    toks.tofile(tmp_path / "train.bin")
    np.save(tmp_path / "validation.npy", toks.astype(np.int32)[::-1].copy())
    ds = TokenFileDataset(tmp_path, sequence_length=64, split="train")
    assert len(ds) == 1000 // 64
    it = ds[3]
    assert it["input_ids"].dtype == torch.int64 and torch.equal(it["input_ids"], torch.from_numpy(toks[192:256].astype(np.int64)))
    dv = TokenFileDataset(tmp_path, 64, "validation")
    assert int(dv[0]["input_ids"][0]) == int(toks[-1])
    assert len(TokenFileDataset(tmp_path / "train.bin", 100)) == 10
    with pytest.raises(FileNotFoundError):
        TokenFileDataset(tmp_path, 64, "test")
    with pytest.raises(ValueError):
        TokenFileDataset(tmp_path / "train.bin", 5000)
    with pytest.raises(IndexError):
        ds[len(ds)]
    big = (np.arange(300, dtype=np.uint32) + 70000)
    big.tofile(tmp_path / "wide.bin")
    (tmp_path / "wide.bin.u32").write_text("")
    assert int(TokenFileDataset(tmp_path / "wide.bin", 100)[1]["input_ids"][0]) == 70100


def test_causal_collator_labels_are_the_inputs():
    items = [{"input_ids": torch.arange(8) + 10 * i} for i in range(3)]
    b = CausalLMCollator()(items)
    assert b["input_ids"].shape == (3, 8) and torch.equal(b["input_ids"], b["labels"]) and b["labels"].data_ptr() != b["input_ids"].data_ptr()


def test_masked_lm_collator_follows_hf_rule():
    V, mask = 50265, 50264
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(3, V - 1, (64, 512), generator=g)
    ids[:, 0], ids[:, -1] = 0, 2
    ids[:, -5:-1] = 1  # pad
    items = [{"input_ids": r} for r in ids]
    col = MaskedLMCollator(V, mask, seed=7)
    b = col(items, step=11)
    sel = b["labels"] != -100
    special = (ids == 0) | (ids == 1) | (ids == 2)
    assert not (sel & special).any(), "special tokens are never selected"
    frac = sel.float().sum() / (~special).float().sum()
    assert abs(frac - 0.15) < 0.01, frac
    assert torch.equal(b["labels"][sel], ids[sel]), "labels hold the original token at the selected positions"
    assert torch.equal(b["input_ids"][~sel], ids[~sel]), "unselected positions are untouched"
    to_mask = (b["input_ids"] == mask) & sel
    kept = (b["input_ids"] == ids) & sel
    n = sel.sum().item()
    assert abs(to_mask.sum().item() / n - 0.8) < 0.02 and abs(kept.sum().item() / n - 0.1) < 0.02
    # a function of (seed, step): a resumed run reproduces the batch; another step draws another mask
    again = MaskedLMCollator(V, mask, seed=7)(items, step=11)
    assert torch.equal(again["input_ids"], b["input_ids"]) and torch.equal(again["labels"], b["labels"])
    assert not torch.equal(col(items, step=12)["labels"], b["labels"])


def test_epoch_sampler_partitions_ranks_and_resumes():
    n, mbs, W = 103, 4, 2
    s = [EpochSampler(n, mbs, W, r, seed=3) for r in range(W)]
    per_epoch = n // (mbs * W)
    seen = torch.cat([s[r].rows(t) for t in range(per_epoch) for r in range(W)])
    assert seen.numel() == per_epoch * mbs * W and seen.unique().numel() == seen.numel(), "one epoch visits every row at most once"
    e1 = torch.cat([s[r].rows(per_epoch + t) for t in range(per_epoch) for r in range(W)])
    assert not torch.equal(e1, seen), "a new permutation per epoch"
    fresh = EpochSampler(n, mbs, W, 1, seed=3)
    assert torch.equal(fresh.rows(per_epoch + 5), s[1].rows(per_epoch + 5)), "stateless in the micro-step index: resumable"
    with pytest.raises(ValueError):
        EpochSampler(5, 4, 2, 0)


def test_training_script_argument_plumbing(tmp_path):
    import training as T

    args = json.loads((ROOT / "tests" / "golden" / "readme_training_arguments.json").read_text())
    zero, fsdp = T._sharding_of(args)
    assert zero == "1" and fsdp == "no_shard"  # the README example is zero_1
    assert T._sharding_of({"deepspeed": None, "fsdp": ["shard_grad_op", "auto_wrap"]}) == ("0", "shard_grad_op")
    assert T._sharding_of({"deepspeed": {"zero_optimization": {"stage": 2}}, "fsdp": ""}) == ("2", "no_shard")
    cls, kw = T.get_optimizer_cls_and_kwargs("pythia-160m", using_deepspeed=True)
    assert cls is torch.optim.Adam and kw["lr"] == 6e-4
    (tmp_path / "checkpoint-2").mkdir()
    (tmp_path / "checkpoint-2" / "trainer_state.json").write_text("{}")
    (tmp_path / "checkpoint-10").mkdir()
    (tmp_path / "checkpoint-10" / "trainer_state.json").write_text("{}")
    (tmp_path / "checkpoint-30").mkdir()  # incomplete: no trainer_state.json
    assert T.latest_checkpoint(tmp_path).name == "checkpoint-10"
    np.arange(4098 * 3, dtype=np.uint16).tofile(tmp_path / "train.bin")
    assert len(T.get_dataset("pythia-70m", tmp_path, "train")) == 6  # 2049-token windows
    assert isinstance(T.get_data_collator("roberta"), MaskedLMCollator) and isinstance(T.get_data_collator("pythia-1b"), CausalLMCollator)
