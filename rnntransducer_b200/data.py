"""Data-side pieces either side of the hot path.

``DistributedBucketSampler`` -- length-bucketed distributed sampler with the constructor, attributes and
index streams of the reference's (datasampler.py:10-99; unused there, ``datamodule.py:12-13``): indices
sorted by length, longest first, padded (or trimmed) to a multiple of the world size, rank-strided.  Rank
r therefore gets the r-th longest, (r+W)-th longest, ... sequence: at every position of the epoch the W
ranks hold neighbours of the sorted order, so their batches have the same shape to within one sequence
and nobody waits at the gradient all-reduce (SURVEY.md 8(e) "efficiency risks").  Added on top (off by
default, so the default stream is the reference's): ``shuffle=True`` permutes whole *global batches*
per epoch (seed + epoch), which keeps both the bucketing and the rank balance.

``collate_sorted`` -- the reference collate (dataloader.py:16-49: the 7-tuple ``input_audios,
audio_lengths, tensor_audio_lengths, input_texts, text_lengths, targets, target_lengths``) with the batch
ordered by audio length, longest first, which lets the bidirectional encoder pack with
``enforce_sorted=True`` (networks.py): no sort, no index upload, no gathers.
"""
from __future__ import annotations

import math
from typing import Iterator, List, Optional, Sequence

import torch
import torch.distributed as dist
from torch.nn.utils.rnn import pad_sequence
from torch.utils.data import Sampler


class DistributedBucketSampler(Sampler):
    def __init__(self, dataset=None, num_replicas: Optional[int] = None, rank: Optional[int] = None,
                 drop_last: bool = False, shuffle: bool = False, lengths: Optional[Sequence[int]] = None,
                 model_input_name: Optional[str] = None, batch_size: int = 1, seed: int = 0):
        if dataset is None and lengths is None:
            raise ValueError("One of dataset and lengths must be provided.")
        if num_replicas is None:
            if not (dist.is_available() and dist.is_initialized()):
                raise RuntimeError("Requires distributed package to be available")
            num_replicas = dist.get_world_size()
        if rank is None:
            if not (dist.is_available() and dist.is_initialized()):
                raise RuntimeError("Requires distributed package to be available")
            rank = dist.get_rank()
        if not 0 <= rank < num_replicas:
            raise ValueError(f"rank {rank} outside [0, {num_replicas})")
        if lengths is None:
            key = model_input_name if model_input_name is not None else "input_ids"
            first = dataset[0]
            if not hasattr(first, "keys") or key not in first:
                raise ValueError("Can only automatically infer lengths for datasets whose items are "
                                 f"dictionaries with an '{key}' key.")
            lengths = [len(item[key]) for item in dataset]
        elif isinstance(lengths, torch.Tensor):
            lengths = lengths.tolist()
        self.lengths: List[int] = [int(n) for n in lengths]
        self.num_replicas, self.rank = int(num_replicas), int(rank)
        self.drop_last, self.shuffle = bool(drop_last), bool(shuffle)
        self.batch_size, self.seed, self.epoch = int(batch_size), int(seed), 0
        n = len(self.lengths)
        if self.drop_last and n % self.num_replicas != 0:
            self.num_samples = math.ceil((n - self.num_replicas) / self.num_replicas)
        else:
            self.num_samples = math.ceil(n / self.num_replicas)
        self.total_size = self.num_samples * self.num_replicas

    def set_epoch(self, epoch: int) -> None:
        self.epoch = int(epoch)

    def get_bucket_indices(self, lengths: Sequence[int]) -> List[int]:
        """Indices by length, longest first; equal lengths keep their dataset order."""
        order = torch.argsort(torch.as_tensor(list(lengths), dtype=torch.int64), descending=True, stable=True)
        return order.tolist()

    def __iter__(self) -> Iterator[int]:
        indices = self.get_bucket_indices(self.lengths)
        if not self.drop_last:
            indices += indices[: self.total_size - len(indices)]  # wrap around: the longest ones again
        else:
            indices = indices[: self.total_size]
        if len(indices) != self.total_size:
            raise RuntimeError("dataset smaller than the padding it needs")
        if self.shuffle:  # permute whole global batches (batch_size sequences on each of the W ranks)
            chunk = self.batch_size * self.num_replicas
            n_chunks = math.ceil(self.total_size / chunk)
            g = torch.Generator().manual_seed(self.seed + self.epoch)
            order = torch.randperm(n_chunks, generator=g).tolist()
            indices = [i for c in order for i in indices[c * chunk:(c + 1) * chunk]]
        return iter(indices[self.rank:self.total_size:self.num_replicas])

    def __len__(self) -> int:
        return self.num_samples


def collate_sorted(batch, pad_token_id: int = 0, n_mels: Optional[int] = None):
    """dataloader.py:16-49 on a list of ``{"input_values": f32 [T_i, n_mels], "input_ids": ints}`` items,
    longest audio first.  Returns the reference's 7-tuple."""
    batch = sorted(batch, key=lambda s: -int(s["input_values"].size(0)))  # stable: ties keep their order
    audios = [s["input_values"] for s in batch]
    audio_lengths = [int(a.size(0)) for a in audios]
    if n_mels is not None and audios[0].size(-1) != n_mels:
        raise ValueError("feature dimension of the data differs from the configured n_mels")
    targets = [torch.as_tensor(s["input_ids"], dtype=torch.int32) for s in batch]
    target_lengths = torch.tensor([len(t) for t in targets], dtype=torch.int32)
    texts = [torch.cat((torch.full((1,), pad_token_id, dtype=torch.int64), t.to(torch.int64))) for t in targets]
    text_lengths = [int(t.numel()) for t in texts]
    return (pad_sequence(audios, batch_first=True, padding_value=pad_token_id), audio_lengths,
            torch.tensor(audio_lengths, dtype=torch.int32),
            pad_sequence(texts, batch_first=True, padding_value=pad_token_id), text_lengths,
            pad_sequence(targets, batch_first=True, padding_value=pad_token_id), target_lengths)
