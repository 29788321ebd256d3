"""CPU restatement of the reference joint + loss step.  TEST INFRASTRUCTURE ONLY.

``joint_concat_gelu`` follows ``/root/reference/networks/transducer.py:54-71`` step by
step (unsqueeze -> repeat x2 -> cat -> GELU(approximate="tanh") -> Linear(2H -> V)),
materialising every intermediate exactly as the reference does -- that is the CPU
baseline the bench times.  ``joint_add_tanh`` is the north_star's alternative joint,
``fc(tanh(enc[:, :, None] + dec[:, None]))`` (semantics of
``torchaudio.models.rnnt._Joiner(activation="tanh")``, torchaudio/models/rnnt.py:392-449);
the reference has no such joint, so that mode's parity is against this statement.

``joint_loss_fwd_bwd`` = joint -> C loss oracle (warp_cpu.c) -> autograd backward through
the joint, i.e. what ``model.py:56-57`` + ``loss.backward()`` do on the CPU validation
path (``model.py:65-74``).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import warp_cpu


def joint_concat_gelu(enc, dec, weight, bias):
    """transducer.py:54-71.  enc [B,T,He], dec [B,U1,Hd], weight [V,He+Hd], bias [V]."""
    if enc.dim() == 3 and dec.dim() == 3:
        input_length = enc.size(1)
        target_length = dec.size(1)
        enc = enc.unsqueeze(2).repeat([1, 1, target_length, 1])
        dec = dec.unsqueeze(1).repeat([1, input_length, 1, 1])
    out = torch.cat((enc, dec), dim=-1)
    out = F.gelu(out, approximate="tanh")
    return F.linear(out, weight, bias)


def joint_add_tanh(enc, dec, weight, bias):
    """fc(tanh(enc_t + dec_u)); weight [V,H]."""
    if enc.dim() == 3 and dec.dim() == 3:
        out = enc.unsqueeze(2) + dec.unsqueeze(1)
    else:
        out = enc + dec
    return F.linear(torch.tanh(out), weight, bias)


JOINTS = {"concat_gelu": joint_concat_gelu, "add_tanh": joint_add_tanh}


def joint_loss_fwd_bwd(enc, dec, weight, bias, labels, act_lens, label_lens, blank=0,
                       reduction="mean", mode="concat_gelu", num_threads=0, dtype=torch.float32):
    """One full CPU step.  Inputs are numpy arrays or CPU tensors.  Returns numpy results:
    costs [B], loss (reduced, shape (1,) for mean/sum), d_enc, d_dec, d_weight, d_bias, logits.
    Gradients are those of the REDUCED loss."""
    t = lambda a: torch.as_tensor(np.asarray(a)).to(dtype).clone().requires_grad_(True)
    enc_t, dec_t, w_t, b_t = t(enc), t(dec), t(weight), t(bias)
    logits = JOINTS[mode](enc_t, dec_t, w_t, b_t)
    np_dtype = np.float32 if dtype == torch.float32 else np.float64
    res = warp_cpu.rnnt_loss_cpu(logits.detach().numpy(), labels, act_lens, label_lens, blank,
                                 want_grad=True, num_threads=num_threads, dtype=np_dtype)
    B = logits.shape[0]
    scale = {"mean": 1.0 / B, "sum": 1.0, "none": 1.0}[reduction]
    logits.backward(torch.from_numpy(res["grads"]).to(dtype) * scale)
    return dict(costs=res["costs"], loss=warp_cpu.reduce_costs(res["costs"], reduction),
                d_enc=enc_t.grad.numpy(), d_dec=dec_t.grad.numpy(), d_weight=w_t.grad.numpy(),
                d_bias=b_t.grad.numpy(), logits=logits.detach().numpy(), threads=res["threads"])


def joint_loss_fwd_bwd_chunked(enc, dec, weight, bias, labels, act_lens, label_lens, blank=0,
                               reduction="mean", mode="concat_gelu", num_threads=0, dtype=torch.float64,
                               frames_per_chunk=None, budget_bytes=1 << 30):
    """The same CPU step as :func:`joint_loss_fwd_bwd` at BASELINE's FULL sizes (cfg 2/3/4).

    The arithmetic is the reference's, cell by cell (``transducer.py:54-71``: repeat -> cat -> GELU ->
    Linear, then the loss, then autograd); only the *evaluation order* differs: the ``[B,T,U1,2H]``
    intermediates (45 GB at cfg 3 in fp32) are formed one (utterance, block of frames) at a time and
    thrown away, the joint is re-evaluated per block for the backward (``logits_block.backward(
    d_logits_block)`` accumulates into the same leaves autograd would reach in one piece).  The dense
    logits ``[B,T,U1,V]`` and their gradient ARE held whole (the loss oracle wants them).
    Returns numpy costs, loss, d_enc, d_dec, d_weight, d_bias (gradients of the REDUCED loss)."""
    t = lambda a: torch.as_tensor(np.asarray(a)).to(dtype).clone().requires_grad_(True)
    enc_t, dec_t, w_t, b_t = t(enc), t(dec), t(weight), t(bias)
    B, T, _ = enc_t.shape
    U1, V = dec_t.shape[1], w_t.shape[0]
    K = w_t.shape[1]
    if frames_per_chunk is None:  # ~4 live copies of the [tc, U1, K] intermediate under autograd
        frames_per_chunk = max(1, int(budget_bytes // (4 * U1 * K * enc_t.element_size())))
    np_dtype = np.float32 if dtype == torch.float32 else np.float64
    logits = np.empty((B, T, U1, V), dtype=np_dtype)
    joint = JOINTS[mode]
    with torch.no_grad():
        for b in range(B):
            for t0 in range(0, T, frames_per_chunk):
                t1 = min(T, t0 + frames_per_chunk)
                logits[b, t0:t1] = joint(enc_t[b:b + 1, t0:t1], dec_t[b:b + 1], w_t, b_t)[0].numpy()
    res = warp_cpu.rnnt_loss_cpu(logits, labels, act_lens, label_lens, blank, want_grad=True,
                                 num_threads=num_threads, dtype=np_dtype)
    del logits
    scale = {"mean": 1.0 / B, "sum": 1.0, "none": 1.0}[reduction]
    grads = res["grads"]
    for b in range(B):
        for t0 in range(0, T, frames_per_chunk):
            t1 = min(T, t0 + frames_per_chunk)
            g = torch.from_numpy(grads[b, t0:t1]).to(dtype)
            if not bool(g.any()):
                continue  # frames past the utterance's length: exactly zero gradient
            out = joint(enc_t[b:b + 1, t0:t1], dec_t[b:b + 1], w_t, b_t)
            out.backward((g * scale).unsqueeze(0))
    z = lambda p: (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
    return dict(costs=res["costs"], loss=warp_cpu.reduce_costs(res["costs"], reduction),
                d_enc=z(enc_t), d_dec=z(dec_t), d_weight=z(w_t), d_bias=z(b_t), threads=res["threads"])
