"""JointNet drop-in (reference networks/transducer.py:27-93) with a fused joint + RNN-T loss.

Same constructor (``JointNet(transnet_params, prednet_params, num_classes)``), same attributes
and state_dict keys (``encoder``, ``decoder``, ``num_classes``, ``act_func``, ``fc``), same
``joint(enc, dec)`` for 3-D and 1-D inputs and the same ``forward(input_audios, audio_lengths,
input_texts, text_lengths)``.  The difference: ``forward`` returns a :class:`JointLogits` handle
instead of the ``[B,T,U+1,V]`` tensor.  ``RNNTLoss`` recognises the handle and runs joint and loss
fused (the logits and the reference's ``[B,T,U+1,2H]`` repeat/cat/GELU intermediates never reach
HBM); anything else that touches the handle gets the dense tensor (``materialize()``).

Modes
  ``concat_gelu`` (default, reference-exact, checkpoint compatible): ``fc(gelu_tanh([e_t ; d_u]))``
      -- evaluated in the factorised form P_enc[t] + P_dec[u].
  ``add_tanh`` (north_star's alternative; ``fc.weight`` is ``[V,H]``): ``fc(tanh(e_t + d_u))``.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import loss as _loss
from .networks import AudioTransNet, TextPredNet

MODES = ("concat_gelu", "add_tanh")


def joint_dense(enc: Tensor, dec: Tensor, weight: Tensor, bias: Optional[Tensor],
                mode: str = "concat_gelu") -> Tensor:
    """Eager dense joint -> ``[B,T,U+1,V]`` (3-D inputs) or ``[V]``/``[..., V]`` (decode form,
    transducer.py:125,309).  Used only when somebody needs the real logits tensor."""
    if mode == "concat_gelu":
        if enc.dim() == 3 and dec.dim() == 3:
            He = enc.size(-1)
            penc = F.linear(F.gelu(enc, approximate="tanh"), weight[:, :He], bias)
            pdec = F.linear(F.gelu(dec, approximate="tanh"), weight[:, He:])
            return penc.unsqueeze(2) + pdec.unsqueeze(1)
        out = torch.cat((enc, dec), dim=-1)
        return F.linear(F.gelu(out, approximate="tanh"), weight, bias)
    if mode == "add_tanh":
        if enc.dim() == 3 and dec.dim() == 3:
            return F.linear(torch.tanh(enc.unsqueeze(2) + dec.unsqueeze(1)), weight, bias)
        return F.linear(torch.tanh(enc + dec), weight, bias)
    raise ValueError(f"mode must be one of {MODES}")


def joint_rnnt_costs(enc: Tensor, dec: Tensor, weight: Tensor, bias: Optional[Tensor],
                     labels: Tensor, act_lens: Tensor, label_lens: Tensor, blank: int = 0,
                     mode: str = "concat_gelu", gemm: str = "fp32",
                     deterministic: bool = False) -> Tensor:
    """Fused joint + RNN-T loss: per-utterance costs ``[B]``, differentiable w.r.t. enc, dec,
    weight, bias.  enc ``[B,T,He]``, dec ``[B,U+1,Hd]``."""
    if enc.device.type != "cuda":
        raise RuntimeError("rnntransducer_b200: inputs must be CUDA tensors (there is no CPU fallback)")
    if enc.dim() != 3 or dec.dim() != 3:
        raise RuntimeError("enc must be [B,T,H] and dec [B,U+1,H]")
    V = weight.shape[0]
    if bias is None:
        bias = weight.new_zeros(V)
    if mode == "concat_gelu":
        He = enc.size(-1)
        if weight.shape[1] != He + dec.size(-1):
            raise RuntimeError("fc.weight must be [V, enc_dim + dec_dim] in concat_gelu mode")
        # two small projections (B*(T+U1) rows instead of B*T*U1); autograd carries them back
        penc, pdec = _loss.project_concat_gelu(enc, dec, weight, bias)
        return _loss._ConcatGeluRNNT.apply(penc, pdec, labels, act_lens, label_lens, int(blank),
                                           bool(deterministic))
    if mode == "add_tanh":
        from .joint_add_tanh import add_tanh_rnnt_costs
        return add_tanh_rnnt_costs(enc, dec, weight, bias, labels, act_lens, label_lens, int(blank),
                                   gemm)
    raise ValueError(f"mode must be one of {MODES}")


def joint_rnnt_loss(enc, dec, weight, bias, labels, act_lens, label_lens, blank=0,
                    reduction="mean", mode="concat_gelu", gemm="fp32", warp_compat=True,
                    deterministic=False):
    """Reduced loss of the fused path (``(1,)`` like warp-transducer, 0-d with ``warp_compat=False``).  In the
    reference's own joint the mean / sum over the utterances is folded into the kernels (one autograd node:
    no reduction kernel, no broadcast multiply in the backward)."""
    if mode == "concat_gelu" and reduction in ("mean", "sum") and enc.is_cuda and enc.dim() == 3 and dec.dim() == 3 \
            and enc.shape[0] > 0 and weight.shape[1] == enc.size(-1) + dec.size(-1):
        if bias is None:
            bias = weight.new_zeros(weight.shape[0])
        penc, pdec = _loss.project_concat_gelu(enc, dec, weight, bias)
        loss, _ = _loss._ConcatGeluRNNTLoss.apply(penc, pdec, labels, act_lens, label_lens, int(blank),
                                                   bool(deterministic), reduction == "mean")
        return loss if warp_compat else loss.reshape(())
    costs = joint_rnnt_costs(enc, dec, weight, bias, labels, act_lens, label_lens, blank, mode, gemm,
                             deterministic)
    return _loss._reduce(costs, reduction, warp_compat)


class JointLogits:
    """Lazy stand-in for ``logits[B,T,U+1,V]`` carrying what the fused path needs."""

    def __init__(self, enc: Tensor, dec: Tensor, weight: Tensor, bias: Optional[Tensor], mode: str,
                 gemm: str = "fp32"):
        self.enc, self.dec, self.weight, self.bias = enc, dec, weight, bias
        self.mode, self.gemm = mode, gemm
        self._dense = None

    # -- what RNNTLoss calls ------------------------------------------------------------------
    def costs(self, labels, act_lens, label_lens, blank=0, deterministic=False):
        return joint_rnnt_costs(self.enc, self.dec, self.weight, self.bias, labels, act_lens,
                                label_lens, blank, self.mode, self.gemm, deterministic)

    def loss(self, labels, act_lens, label_lens, blank=0, reduction="mean", warp_compat=True, deterministic=False):
        return joint_rnnt_loss(self.enc, self.dec, self.weight, self.bias, labels, act_lens, label_lens, blank,
                               reduction, self.mode, self.gemm, warp_compat, deterministic)

    # -- tensor-like surface ------------------------------------------------------------------
    @property
    def shape(self):
        return torch.Size((self.enc.shape[0], self.enc.shape[1], self.dec.shape[1],
                           self.weight.shape[0]))

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def dim(self):
        return 4

    @property
    def device(self):
        return self.enc.device

    @property
    def dtype(self):
        return self.enc.dtype

    def materialize(self) -> Tensor:
        """The real ``[B,T,U+1,V]`` tensor (eager; costs B*T*U1*V*4 bytes of HBM)."""
        if self._dense is None:
            self._dense = joint_dense(self.enc, self.dec, self.weight, self.bias, self.mode)
        return self._dense

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        conv = lambda a: a.materialize() if isinstance(a, JointLogits) else a
        args = tuple(conv(a) for a in args)
        kwargs = {k: conv(v) for k, v in (kwargs or {}).items()}
        return func(*args, **kwargs)

    def __getattr__(self, name):  # tensor methods (argmax, cpu, ...) act on the dense tensor
        if name.startswith("__"):
            raise AttributeError(name)
        return getattr(self.materialize(), name)

    def __getitem__(self, idx):
        return self.materialize()[idx]


class JointNet(nn.Module):
    """See module docstring.  ``fused=False`` restores the reference behaviour of returning the
    dense logits from ``forward``."""

    def __init__(self, transnet_params: dict, prednet_params: dict, num_classes: int,
                 mode: str = "concat_gelu", fused: bool = True, gemm: str = "fp32"):
        super().__init__()
        if mode not in MODES:
            raise ValueError(f"mode must be one of {MODES}")
        self.encoder = AudioTransNet(**transnet_params)
        self.decoder = TextPredNet(**prednet_params)
        self.num_classes = num_classes
        self.mode, self.fused, self.gemm = mode, fused, gemm
        if mode == "concat_gelu":
            self.act_func = nn.GELU(approximate="tanh")
            self.fc = nn.Linear(transnet_params["output_size"] + prednet_params["output_size"],
                                num_classes)
        else:
            if transnet_params["output_size"] != prednet_params["output_size"]:
                raise ValueError("add_tanh needs equal encoder / predictor output sizes")
            self.act_func = nn.Tanh()
            self.fc = nn.Linear(transnet_params["output_size"], num_classes)

    def joint(self, encoder_outputs: Tensor, decoder_outputs: Tensor) -> Tensor:
        """Dense joint, 3-D and 1-D forms (transducer.py:41-71)."""
        return joint_dense(encoder_outputs, decoder_outputs, self.fc.weight, self.fc.bias, self.mode)

    def joint_lazy(self, encoder_outputs: Tensor, decoder_outputs: Tensor) -> JointLogits:
        return JointLogits(encoder_outputs, decoder_outputs, self.fc.weight, self.fc.bias, self.mode,
                           self.gemm)

    def forward(self, input_audios: Tensor, audio_lengths, input_texts: Tensor, text_lengths):
        """transducer.py:73-93.  ``audio_lengths`` / ``text_lengths`` are the collate's host lists."""
        enc_state = self.encoder(input_audios, audio_lengths)
        dec_state, _ = self.decoder(input_texts, text_lengths)
        if self.fused and enc_state.is_cuda:
            return self.joint_lazy(enc_state, dec_state)
        return self.joint(enc_state, dec_state)

    @torch.no_grad()
    def recognize_greedy(self, inputs: Tensor, inputs_lengths, blank_token_id: int, max_iters: int = 3,
                         respect_lengths: bool = False) -> Tensor:
        """Greedy decode of the whole batch at once (reference transducer.py:95-145, called from every
        validation step, model.py:76).

        The reference walks utterance by utterance, frame by frame, symbol by symbol, with an ``.item()``
        (a device sync) per symbol.  Here all B utterances advance together: per (frame, iteration) one
        joint evaluation ``[B, V]``, one arg-max, one predictor step for the whole batch whose new state is
        kept only in the rows that emitted a symbol -- same tokens, no host round trip until the single
        read of the longest hypothesis' length at the end.  In the reference's own joint the encoder half
        of the logits, ``gelu(enc) W_e^T + b`` for all B*T frames, is computed ONCE by the fused path's
        projection kernel (``rnntb200_joint_cg_project``); a step then costs a ``[B, Hd] x [Hd, V]``
        product.

        Like the reference: every padded frame is decoded (``respect_lengths=True`` stops each row at its
        own length instead), at most ``max_iters`` symbols per frame, a symbol equal to the previously
        *kept* one is fed to the predictor but not kept.  Returns ``LongTensor [B, L]`` on the inputs'
        device, L = longest hypothesis, shorter rows padded with ``blank_token_id`` (the reference stacks
        the rows and therefore only works when all hypotheses are equally long; then the two agree)."""
        enc = self.encoder(inputs, inputs_lengths)
        B, T, He = enc.shape
        dev = enc.device
        W, bias = self.fc.weight, self.fc.bias
        tok = torch.full((B, 1), blank_token_id, dtype=torch.long, device=dev)
        dec_out, hidden = self.decoder(tok)
        dec_out = dec_out[:, 0]
        factorised = self.mode == "concat_gelu"
        if factorised:  # logits(t) = P_enc[:, t] + gelu(dec) W_d^T
            penc, _ = _loss.project_concat_gelu(enc, dec_out.unsqueeze(1), W, bias)
            W_d = W[:, He:]
        lens = None
        if respect_lengths:
            lens = torch.tensor([int(n) for n in inputs_lengths], device=dev)
        rows = torch.arange(B, device=dev)
        kept = torch.full((B, T * max_iters + 1), blank_token_id, dtype=torch.long, device=dev)
        count = torch.zeros(B, dtype=torch.long, device=dev)
        last = torch.full((B,), blank_token_id, dtype=torch.long, device=dev)
        sel = lambda m, new, old: torch.where(m.view((1, B, 1)), new, old)
        for t in range(T):
            active = torch.ones(B, dtype=torch.bool, device=dev) if lens is None else lens > t
            for _ in range(max_iters):
                if factorised:
                    logits = penc[:, t] + F.linear(F.gelu(dec_out, approximate="tanh"), W_d)
                else:
                    logits = joint_dense(enc[:, t], dec_out, W, bias, self.mode)
                pred = logits.argmax(dim=-1)
                emit = active & (pred != blank_token_id)
                new_out, new_hidden = self.decoder(pred.unsqueeze(1), None, hidden)
                dec_out = torch.where(emit.unsqueeze(1), new_out[:, 0], dec_out)
                if isinstance(hidden, tuple):
                    hidden = tuple(sel(emit, n, o) for n, o in zip(new_hidden, hidden))
                else:
                    hidden = sel(emit, new_hidden, hidden)
                keep = emit & (pred != last)
                kept[rows, count] = torch.where(keep, pred, kept[rows, count])
                count = count + keep
                last = torch.where(keep, pred, last)
                active = emit
        return kept[:, :max(int(count.max()), 0)]  # the one device sync of the decode
