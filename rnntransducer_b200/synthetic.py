"""Seeded synthetic batches in the shapes BASELINE.json names.

There is no dataset on the GPU box, so bench and tests draw encoder/predictor outputs
``enc [B,T,H]``, ``dec [B,U+1,H]`` ~ N(0,1), ``fc`` parameters with torch's default
``nn.Linear`` init range U(-1/sqrt(K), 1/sqrt(K)), labels ~ UniformInt[1, V-1] (blank = 0 is
never a label: 72 graphemes + blank) and either *full* or *ragged* lengths.  Everything is
generated on the CPU generator so the same seed gives the same batch on every device.
"""
from __future__ import annotations

import math

import torch

# BASELINE.json:configs (H = 512 assumed where unstated)
CONFIGS = {
    1: dict(name="cfg1 B4 T100 U20 V73 H320", B=4, T=100, U=20, V=73, H=320),
    2: dict(name="cfg2 KsponSpeech-shaped B32 T400 U80 V73 H512", B=32, T=400, U=80, V=73, H=512),
    3: dict(name="cfg3 long-utterance B8 T1500 U300 V73 H512", B=8, T=1500, U=300, V=73, H=512),
    4: dict(name="cfg4 large-vocab B16 T400 U100 V1024 H512", B=16, T=400, U=100, V=1024, H=512),
}


def make_lengths(B, T, U, ragged, gen):
    if not ragged:
        return (torch.full((B,), T, dtype=torch.int32), torch.full((B,), U, dtype=torch.int32))
    act = torch.randint(max(T // 2, 1), T + 1, (B,), generator=gen, dtype=torch.int32)
    lab = torch.randint(U // 2, U + 1, (B,), generator=gen, dtype=torch.int32)
    act[0], lab[0] = T, U  # one utterance spans the padded box
    return act, lab


def make_batch(B, T, U, V, H, mode="concat_gelu", ragged=False, seed=1234, device="cpu",
               dtype=torch.float32, **_unused):
    """Returns dict(enc, dec, weight, bias, labels, act_lens, label_lens)."""
    gen = torch.Generator(device="cpu").manual_seed(int(seed))
    enc = torch.randn(B, T, H, generator=gen)
    dec = torch.randn(B, U + 1, H, generator=gen)
    K = 2 * H if mode == "concat_gelu" else H
    bound = 1.0 / math.sqrt(K)
    weight = (torch.rand(V, K, generator=gen) * 2 - 1) * bound
    bias = (torch.rand(V, generator=gen) * 2 - 1) * bound
    labels = torch.randint(1, V, (B, max(U, 1)), generator=gen, dtype=torch.int32)[:, :U]
    act_lens, label_lens = make_lengths(B, T, U, ragged, gen)
    if ragged:  # zero the label padding, as the reference collate does (dataloader.py:41-43)
        mask = torch.arange(U)[None, :] >= label_lens[:, None]
        labels = labels.masked_fill(mask, 0)
    out = dict(enc=enc.to(dtype), dec=dec.to(dtype), weight=weight.to(dtype), bias=bias.to(dtype),
               labels=labels.contiguous(), act_lens=act_lens, label_lens=label_lens)
    return {k: v.to(device) for k, v in out.items()}


def make_dense_logits(B, T, U, V, ragged=False, seed=1234, device="cpu", dtype=torch.float32):
    """Dense-logits batch for the plain RNNTLoss drop-in: logits ~ N(0,1)."""
    gen = torch.Generator(device="cpu").manual_seed(int(seed))
    logits = torch.randn(B, T, U + 1, V, generator=gen)
    labels = torch.randint(1, V, (B, max(U, 1)), generator=gen, dtype=torch.int32)[:, :U]
    act_lens, label_lens = make_lengths(B, T, U, ragged, gen)
    out = dict(logits=logits.to(dtype), labels=labels.contiguous(), act_lens=act_lens,
               label_lens=label_lens)
    return {k: v.to(device) for k, v in out.items()}


def count_cells(act_lens, label_lens) -> int:
    """cells = sum_b T_b * (U_b + 1) -- the unit of BASELINE.json's metric."""
    return int((act_lens.long() * (label_lens.long() + 1)).sum().item())
