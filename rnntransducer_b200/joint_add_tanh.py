"""Fused joint + RNN-T loss, ``add_tanh`` mode: ``logits(t,u,:) = fc(tanh(enc_t + dec_u))``.

This is the north_star's joint (broadcast-add, tanh, Linear H -> V) -- the one dense contraction
per lattice cell.  The reference has no such joint (its own is concat -> GELU -> Linear 2H -> V,
networks/transducer.py:64-69, served by ``joint_cg``); semantics follow
``torchaudio.models.rnnt._Joiner(activation="tanh")``.

``gemm`` selects the arithmetic of the contraction (include/rnnt_b200.h ``rnntb200_gemm_t``):
  ``"fp32"``    CUDA-core FFMA, fp32 parity tolerance (loss 1e-5 rel, grads 1e-4 abs);
  ``"bf16"``    tcgen05 ``kind::f16`` MMA, bf16 operands, fp32 accumulation in TMEM -- looser,
                separately stated tolerance (tests/test_gpu_joint_at.py);
  ``"tf32x3"``  reserved.
"""
from __future__ import annotations

import torch

from . import _lib
from .loss import _contiguous, _ptr, _stream, _validate

GEMMS = {"fp32": 0, "bf16": 1, "tf32x3": 2}


class _AddTanhRNNT(torch.autograd.Function):
    @staticmethod
    def forward(ctx, enc, dec, weight, bias, labels, act_lens, label_lens, blank, gemm):
        B, T, H = enc.shape
        U1 = dec.shape[1]
        V = weight.shape[0]
        if dec.shape[2] != H or weight.shape[1] != H:
            raise RuntimeError("add_tanh joint needs enc [B,T,H], dec [B,U+1,H], fc.weight [V,H]")
        _validate((B, T, U1, V), enc.device, labels, act_lens, label_lens, blank)
        enc, dec = enc.contiguous().float(), dec.contiguous().float()
        weight, bias = weight.contiguous().float(), bias.contiguous().float()
        labels, act_lens, label_lens = _contiguous(labels, act_lens, label_lens)
        f32 = dict(device=enc.device, dtype=torch.float32)
        costs = torch.empty(B, **f32)
        lp2 = torch.empty(B, T, U1, 2, **f32)
        lse = torch.empty(B, T, U1, **f32)
        alpha, beta = (torch.empty(B, T, U1, device=enc.device, dtype=torch.int32) for _ in range(2))
        lib = _lib.load()
        ws_bytes = lib.rnntb200_joint_at_workspace_bytes(V, H, gemm)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=enc.device) if ws_bytes else None
        with torch.cuda.device(enc.device):
            _lib.check(lib.rnntb200_joint_at_fwd(
                _ptr(enc), _ptr(dec), _ptr(weight), _ptr(bias), gemm, _ptr(labels), _ptr(act_lens),
                _ptr(label_lens), B, T, U1, V, H, blank, _ptr(costs), _ptr(lp2), _ptr(lse),
                _ptr(alpha), _ptr(beta), _ptr(ws), ws_bytes, _stream()), "rnntb200_joint_at_fwd")
        ctx.save_for_backward(enc, dec, weight, bias, labels, act_lens, label_lens, lp2, lse, alpha,
                              beta)
        ctx.blank, ctx.gemm = blank, gemm
        return costs

    @staticmethod
    def backward(ctx, grad_costs):
        (enc, dec, weight, bias, labels, act_lens, label_lens, lp2, lse, alpha,
         beta) = ctx.saved_tensors
        B, T, H = enc.shape
        U1 = dec.shape[1]
        V = weight.shape[0]
        grad_costs = grad_costs.contiguous().to(torch.float32)
        d_enc, d_dec = torch.empty_like(enc), torch.empty_like(dec)
        d_w, d_b = torch.empty_like(weight), torch.empty_like(bias)
        lib = _lib.load()
        ws_bytes = lib.rnntb200_joint_at_workspace_bytes(V, H, ctx.gemm)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=enc.device) if ws_bytes else None
        with torch.cuda.device(enc.device):
            _lib.check(lib.rnntb200_joint_at_bwd(
                _ptr(enc), _ptr(dec), _ptr(weight), _ptr(bias), ctx.gemm, _ptr(labels),
                _ptr(act_lens), _ptr(label_lens), B, T, U1, V, H, ctx.blank, _ptr(lp2), _ptr(lse),
                _ptr(alpha), _ptr(beta), _ptr(grad_costs), _ptr(d_enc), _ptr(d_dec), _ptr(d_w),
                _ptr(d_b), _ptr(ws), ws_bytes, _stream()), "rnntb200_joint_at_bwd")
        return d_enc, d_dec, d_w, d_b, None, None, None, None, None


def add_tanh_rnnt_costs(enc, dec, weight, bias, labels, act_lens, label_lens, blank=0, gemm="fp32"):
    if gemm not in GEMMS:
        raise ValueError(f"gemm must be one of {tuple(GEMMS)}")
    return _AddTanhRNNT.apply(enc, dec, weight, bias, labels, act_lens, label_lens, int(blank),
                              GEMMS[gemm])
