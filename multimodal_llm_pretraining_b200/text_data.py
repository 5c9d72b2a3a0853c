"""Text pretraining data for `scripts/training.py`. The reference wires `scripts/training.py:19-71` only for
its multimodal families and raises NotImplementedError for Pythia / RoBERTa; this is the missing text-LM branch, shaped like the
reference's other branches: a map-style Dataset of token windows + a collator that builds {"input_ids", "labels"} batches
(the keys `src/benchmarking/data.py:17-21` uses and HF Trainer passes to model.forward).

Host-side only (numpy memmap + torch CPU ops feeding pinned batches); nothing here is on the GPU hot path.

  TokenFileDataset   pre-tokenized corpus as ONE flat token file, the format GPT-NeoX / Pythia training consumes after
                     tokenization: `<name>.bin` (raw little-endian uint16, or uint32 with `<name>.bin.u32` next to it / dtype=) or
                     `<name>.npy`. Item i = tokens [i*S, (i+1)*S) as int64. `split` selects the file `<data_path>/<split>.bin|npy`
                     when data_path is a directory.
  CausalLMCollator   labels = input_ids (HF DataCollatorForLanguageModeling(mlm=False) without padding; the model shifts).
  MaskedLMCollator   HF DataCollatorForLanguageModeling(mlm=True).torch_mask_tokens: 15 % of the non-special positions are
                     selected; 80 % of them -> <mask>, 10 % -> a random token, 10 % unchanged; labels = -100 elsewhere.
  EpochSampler       deterministic, resumable order: micro-batch t of rank r = rows perm_epoch[(t*W + r)*mbs : ...], a fresh seeded
                     permutation per epoch; a resumed run restores t and therefore continues exactly where it stopped.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch
from torch.utils.data import Dataset


class TokenFileDataset(Dataset):
    def __init__(self, data_path: str | Path, sequence_length: int, split: str = "train", dtype: str | None = None):
        p = Path(data_path)
        if p.is_dir():
            cands = [p / f"{split}.bin", p / f"{split}.npy"]
            found = [c for c in cands if c.exists()]
            if not found:
                raise FileNotFoundError(f"no {split}.bin / {split}.npy under {p}")
            p = found[0]
        if not p.exists():
            raise FileNotFoundError(p)
        if p.suffix == ".npy":
            self.tokens = np.load(p, mmap_mode="r")
        else:
            if dtype is None:
                dtype = "uint32" if Path(str(p) + ".u32").exists() else "uint16"
            self.tokens = np.memmap(p, dtype=np.dtype(dtype), mode="r")
        if self.tokens.ndim != 1:
            self.tokens = self.tokens.reshape(-1)
        self.S = int(sequence_length)
        self.n = len(self.tokens) // self.S
        if self.n == 0:
            raise ValueError(f"{p}: {len(self.tokens)} tokens are fewer than one sequence of {self.S}")

    def __len__(self) -> int:
        return self.n

    def __getitem__(self, i: int) -> dict[str, torch.Tensor]:
        if not 0 <= i < self.n:
            raise IndexError(i)
        w = np.asarray(self.tokens[i * self.S:(i + 1) * self.S]).astype(np.int64)
        return {"input_ids": torch.from_numpy(w)}


class CausalLMCollator:
    def __call__(self, items: list[dict[str, torch.Tensor]], step: int | None = None) -> dict[str, torch.Tensor]:
        ids = torch.stack([it["input_ids"] for it in items])
        return {"input_ids": ids, "labels": ids.clone()}


class MaskedLMCollator:
    def __init__(self, vocab_size: int, mask_token_id: int, special_token_ids: tuple[int, ...] = (0, 1, 2), mlm_probability: float = 0.15,
                 seed: int = 0):
        self.V, self.mask_id, self.special, self.p = vocab_size, mask_token_id, tuple(special_token_ids), mlm_probability
        self.seed = seed
        self.gen = torch.Generator().manual_seed(seed)

    def __call__(self, items: list[dict[str, torch.Tensor]], step: int | None = None) -> dict[str, torch.Tensor]:
        if step is not None:  # masks as a function of (seed, global micro-batch index): a resumed run draws the same ones
            self.gen.manual_seed(self.seed * 1_000_003 + step)
        inputs = torch.stack([it["input_ids"] for it in items]).clone()
        labels = inputs.clone()
        prob = torch.full(labels.shape, self.p)
        special = torch.zeros_like(labels, dtype=torch.bool)
        for t in self.special + (self.mask_id,):
            special |= labels == t
        prob.masked_fill_(special, 0.0)
        masked = torch.bernoulli(prob, generator=self.gen).bool()
        labels[~masked] = -100
        replaced = torch.bernoulli(torch.full(labels.shape, 0.8), generator=self.gen).bool() & masked
        inputs[replaced] = self.mask_id
        rand = torch.bernoulli(torch.full(labels.shape, 0.5), generator=self.gen).bool() & masked & ~replaced
        words = torch.randint(self.V, labels.shape, dtype=torch.long, generator=self.gen)
        inputs[rand] = words[rand]
        return {"input_ids": inputs, "labels": labels}


class EpochSampler:
    def __init__(self, num_samples: int, micro_batch_size: int, world_size: int = 1, rank: int = 0, seed: int = 0):
        self.n, self.mbs, self.W, self.rank, self.seed = num_samples, micro_batch_size, world_size, rank, seed
        self.per_epoch = num_samples // (micro_batch_size * world_size)  # micro-steps per epoch (incomplete tail dropped)
        if self.per_epoch == 0:
            raise ValueError(f"{num_samples} samples do not fill one micro-batch of {micro_batch_size} on {world_size} rank(s)")
        self._epoch, self._perm = -1, None

    def rows(self, micro_step: int) -> torch.Tensor:
        epoch, t = divmod(micro_step, self.per_epoch)
        if epoch != self._epoch:
            self._perm = torch.randperm(self.n, generator=torch.Generator().manual_seed(self.seed * 1_000_003 + epoch))
            self._epoch = epoch
        b = t * self.W + self.rank
        return self._perm[b * self.mbs:(b + 1) * self.mbs]


def get_text_dataset(model_class, data_path, data_split: str) -> TokenFileDataset:
    return TokenFileDataset(data_path, model_class.sequence_length, data_split)


def get_text_collator(model_type: str, model_class, seed: int = 0):
    if model_type == "roberta":
        # roberta-large vocabulary: <s> 0, <pad> 1, </s> 2, <mask> 50264 (HF: RobertaTokenizer)
        return MaskedLMCollator(model_class.vocab_size, mask_token_id=model_class.vocab_size - 1, special_token_ids=(0, 1, 2), seed=seed)
    return CausalLMCollator()
