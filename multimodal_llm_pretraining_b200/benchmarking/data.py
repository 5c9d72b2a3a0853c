"""Synthetic data of the benchmarked path — mirrors src/benchmarking/data.py:8-21 — plus the rank/micro-batch sharding
that HF Trainer + accelerate perform in the reference, stated explicitly so it can be tested bit-exact."""
from __future__ import annotations

import torch
from torch.utils.data import Dataset


class DummyTextModelingDataset(Dataset):
    """`input_ids = randint(0, vocab, (num_samples, seq_len))`, `labels` = a copy; items are dicts of two int64 [S]
    (src/benchmarking/data.py:8-21). The reference is unseeded; `seed` makes parity runs reproducible."""

    def __init__(self, vocab_size: int, sequence_length: int, num_samples: int = 50_000, seed: int | None = None) -> None:
        super().__init__()
        g = None
        if seed is not None:
            g = torch.Generator().manual_seed(seed)
        self.input_ids = torch.randint(0, vocab_size, (num_samples, sequence_length), generator=g)
        self.labels = self.input_ids.clone()

    def __len__(self):
        return len(self.input_ids)

    def __getitem__(self, index):
        return {"input_ids": self.input_ids[index], "labels": self.labels[index]}


def shard_rows(num_samples: int, micro_batch_size: int, world_size: int, rank: int, micro_step: int,
               permutation: torch.Tensor | None = None) -> torch.Tensor:
    """Dataset rows consumed by `rank` at local micro-step `micro_step`: batches of `micro_batch_size` consecutive
    entries of the (optionally permuted) index list are dealt round-robin to ranks — batch index t*W + r — wrapping
    around at the end of the data (accelerate's BatchSamplerShard, as driven by HF Trainer.get_train_dataloader)."""
    batches_total = num_samples // micro_batch_size
    b = (micro_step * world_size + rank) % batches_total
    idx = torch.arange(b * micro_batch_size, (b + 1) * micro_batch_size)
    return permutation[idx] if permutation is not None else idx


class ShardedBatchIterator:
    """Infinite iterator of {"input_ids", "labels"} micro-batches for one rank (pinned host memory when available)."""

    def __init__(self, dataset: DummyTextModelingDataset, micro_batch_size: int, world_size: int = 1, rank: int = 0,
                 shuffle_seed: int | None = 0, pin: bool = True):
        self.ds, self.mbs, self.W, self.rank = dataset, micro_batch_size, world_size, rank
        self.perm = None
        if shuffle_seed is not None:
            self.perm = torch.randperm(len(dataset), generator=torch.Generator().manual_seed(shuffle_seed))
        self.t = 0
        self.pin = pin and torch.cuda.is_available()

    def __iter__(self):
        return self

    def __next__(self) -> dict[str, torch.Tensor]:
        rows = shard_rows(len(self.ds), self.mbs, self.W, self.rank, self.t, self.perm)
        self.t += 1
        ids = self.ds.input_ids[rows]
        lab = self.ds.labels[rows]
        if self.pin:
            ids, lab = ids.pin_memory(), lab.pin_memory()
        return {"input_ids": ids, "labels": lab}
