"""The reference's training step without Lightning (reference model.py:10-60, 110-126; train.py:45-48).

``RNNTransducerStep`` is what ``RNNTransducer.training_step`` computes -- ``logits = jointnet(audios,
audio_lengths, texts, text_lengths)``; ``loss = rnnt_loss(logits, targets, tensor_audio_lengths,
target_lengths)`` -- as one ``nn.Module`` whose ``forward`` returns the loss tensor.  That is the object
to hand to ``torch.nn.parallel.DistributedDataParallel``: DDP hooks the autograd graph of the tensor its
module returns, and the fused ``JointNet.forward`` returns a lazy ``JointLogits`` handle, not a tensor, so
wrapping ``JointNet`` alone would leave DDP without an output to trace.  Attribute names
(``jointnet``, ``rnnt_loss``) are the reference's, so its checkpoints' ``jointnet.*`` keys load.

``configure_optimizers`` restates model.py:110-126: AdamW over all parameters + OneCycleLR stepped per
optimizer step.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn as nn
from torch import Tensor

from .joint import JointNet
from .loss import RNNTLoss


class RNNTransducerStep(nn.Module):
    def __init__(self, prednet_params: dict, transnet_params: dict, jointnet_params: dict,
                 blank_token_id: int = 0, reduction: str = "mean", mode: str = "concat_gelu",
                 gemm: str = "fp32", deterministic: bool = False):
        super().__init__()
        prednet_params = dict(prednet_params)
        prednet_params.setdefault("pad_token_id", blank_token_id)  # model.py:26
        self.blank_token_id = blank_token_id
        self.jointnet = JointNet(dict(transnet_params), prednet_params, mode=mode, gemm=gemm, **jointnet_params)
        self.rnnt_loss = RNNTLoss(blank=blank_token_id, reduction=reduction, deterministic=deterministic)

    def forward(self, input_audios: Tensor, audio_lengths: Sequence[int], tensor_audio_lengths: Tensor,
                input_texts: Tensor, text_lengths: Sequence[int], targets: Tensor,
                target_lengths: Tensor) -> Tensor:
        """The collate's 7-tuple (dataloader.py:49) -> loss, shape ``(1,)`` like warp-transducer's."""
        logits = self.jointnet(input_audios, audio_lengths, input_texts, text_lengths)  # model.py:56
        return self.rnnt_loss(logits, targets, tensor_audio_lengths, target_lengths)    # model.py:57


def configure_optimizers(module: nn.Module, learning_rate: float, weight_decay: float, total_steps: int,
                         warmup_ratio: float = 0.3, final_div_factor: float = 1e4,
                         max_lr: Optional[float] = None):
    """AdamW + OneCycleLR as in model.py:110-126 (the scheduler is stepped once per optimizer step)."""
    optimizer = torch.optim.AdamW([{"params": [p for p in module.parameters()], "name": "OneCycleLR"}],
                                  lr=learning_rate, weight_decay=weight_decay)
    scheduler = torch.optim.lr_scheduler.OneCycleLR(
        optimizer, max_lr=max_lr if max_lr is not None else learning_rate, total_steps=total_steps,
        pct_start=warmup_ratio, final_div_factor=final_div_factor)
    return optimizer, scheduler


def synthetic_training_batch(B: int, T: int, U: int, n_mels: int, vocab: int, ragged: bool = False,
                             seed: int = 0, blank: int = 0):
    """A collate-shaped batch (dataloader.py:16-49) of random log-mel frames and labels on the host:
    (input_audios f32 [B,T,n_mels], audio_lengths list, tensor_audio_lengths i32 [B], input_texts i64
    [B,U+1] = [blank] + labels, text_lengths list, targets i32 [B,U], target_lengths i32 [B])."""
    from .synthetic import make_lengths
    gen = torch.Generator(device="cpu").manual_seed(int(seed))
    audios = torch.randn(B, T, n_mels, generator=gen)
    labels = torch.randint(1, vocab, (B, max(U, 1)), generator=gen, dtype=torch.int32)[:, :U]
    act_lens, label_lens = make_lengths(B, T, U, ragged, gen)
    mask = torch.arange(U)[None, :] >= label_lens[:, None]
    labels = labels.masked_fill(mask, blank)                      # zero-padded targets (dataloader.py:43)
    audios = audios * (torch.arange(T)[None, :, None] < act_lens[:, None, None])  # padded frames = 0 (:41)
    texts = torch.cat((torch.full((B, 1), blank, dtype=torch.int64), labels.to(torch.int64)), 1)
    return (audios, act_lens.tolist(), act_lens.clone(), texts, (label_lens + 1).tolist(), labels.contiguous(),
            label_lens.clone())
