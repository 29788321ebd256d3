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
