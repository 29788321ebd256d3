"""Utterance sharding for the multi-GPU path.

The fused joint + loss exchanges nothing between utterances, so the N-GPU path is: one process per
GPU (torchrun), the batch split by utterance, each rank runs the single-GPU kernels on its shard
with ``reduction="mean"`` and the only collective is the gradient all-reduce DDP performs
(reference train.py:45-48).  With equal shard sizes the mean over ranks of the per-rank means is
the global mean, so results match the single-process run on the concatenated batch.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_indices(batch_size: int, rank: int, world_size: int) -> torch.Tensor:
    """Rank-strided utterance indices -- what ``DistributedSampler(shuffle=False)`` hands rank
    ``rank`` (Lightning injects that sampler for the reference, README.md:52-54)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    return torch.arange(rank, batch_size, world_size)


def shard_batch(batch: dict, rank: int, world_size: int) -> dict:
    """Slice every per-utterance tensor of ``batch`` (first dim == B); parameters are replicated."""
    B = batch["act_lens"].shape[0]
    idx = shard_indices(B, rank, world_size)
    per_utt = ("enc", "dec", "logits", "labels", "act_lens", "label_lens")
    return {k: (v.index_select(0, idx.to(v.device)).contiguous() if k in per_utt else v)
            for k, v in batch.items()}


def allreduce_mean_(tensors, group=None) -> None:
    """In-place mean of gradient tensors over ranks (what DDP's bucket all-reduce computes)."""
    world = dist.get_world_size(group)
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t.div_(world)
