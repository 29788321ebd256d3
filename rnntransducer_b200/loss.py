"""RNNTLoss drop-in (reference call sites: model.py:31,39 construct; model.py:57,74 call).

    loss_fn = RNNTLoss(blank=0, reduction="mean")
    loss = loss_fn(acts, labels, act_lens, label_lens)

``acts`` is either a dense ``[B,T,U+1,V]`` logits tensor (the reference's own calling convention)
or the lazy :class:`~rnntransducer_b200.joint.JointLogits` handle returned by our ``JointNet``; in
the second case joint and loss run fused and the logits are never materialised.

The functional spelling the north_star names -- ``rnnt_loss(acts, labels, act_lens, label_lens,
blank, reduction)`` -- is warp-transducer's ``_RNNT.apply`` argument order.

Everything runs through the C ABI of ``librnnt_b200.so`` (ctypes, raw device pointers, the
caller's stream).  CPU tensors are rejected: there is no CPU fallback.
"""
from __future__ import annotations

import torch

from . import _lib

_DTYPES = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _validate(acts_shape, acts_device, labels, act_lens, label_lens, blank, require_cuda=True):
    """Host-side checks that need no device sync (mirrors the oracle's messages)."""
    if require_cuda and acts_device.type != "cuda":
        raise RuntimeError("rnntransducer_b200: inputs must be CUDA tensors (there is no CPU fallback)")
    B, T, U1, V = acts_shape
    if labels.dtype != torch.int32:
        raise RuntimeError("labels must be int32")
    if act_lens.dtype != torch.int32 or label_lens.dtype != torch.int32:
        raise RuntimeError("act_lens and label_lens must be int32")
    for name, t in (("labels", labels), ("act_lens", act_lens), ("label_lens", label_lens)):
        if t.device != acts_device:
            raise RuntimeError(f"{name} must be on the same device as acts ({acts_device})")
    if labels.dim() != 2 or labels.shape[0] != B:
        raise RuntimeError("labels must be [B, U]")
    if labels.shape[1] + 1 != U1:
        raise RuntimeError(f"acts dim 2 must be labels.shape[1] + 1 (got {U1} vs {labels.shape[1]} + 1)")
    if act_lens.shape != (B,) or label_lens.shape != (B,):
        raise RuntimeError("act_lens and label_lens must be [B]")
    if not 0 <= blank < V:
        raise RuntimeError(f"blank must be in [0, {V}), got {blank}")
    if T <= 0:
        raise RuntimeError("acts must have at least one frame")
    if U1 > 1024:
        raise RuntimeError(f"label sequences longer than 1023 are not supported (U+1 = {U1} > 1024)")


def _contiguous(*tensors):
    """The kernels index labels / act_lens / label_lens with stride 1 from the raw pointer: a strided
    view (``lens[:, 0]``, ``lens[::2]``) must be compacted first (no-op for contiguous tensors)."""
    return tuple(t.contiguous() for t in tensors)


def check_lengths(act_lens, label_lens, T, U1, labels=None, V=None):
    """Optional debug check (DEVICE SYNC): lengths inside the padded box and -- when ``labels`` and
    ``V`` are given -- every label inside the vocabulary.  Without it the kernels clamp such values
    (csrc/common.cuh len_T / len_U / label_at): defined behaviour, wrong training data."""
    if int(act_lens.max()) > T or int(act_lens.min()) < 1:
        raise RuntimeError("act_lens out of range")
    if int(label_lens.max()) + 1 > U1 or int(label_lens.min()) < 0:
        raise RuntimeError("label_lens out of range")
    if labels is not None and V is not None and labels.numel() > 0:
        if int(labels.max()) >= V or int(labels.min()) < 0:
            raise RuntimeError(f"labels must be in [0, {V})")


class _DenseRNNT(torch.autograd.Function):
    """costs[B] = -log P(y|x) from dense logits (K5 front-end + alpha/beta sweeps; K4 gradient)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, acts, labels, act_lens, label_lens, blank):
        if acts.dtype not in _DTYPES:
            raise RuntimeError("acts must be float32, float16 or bfloat16")
        _validate(acts.shape, acts.device, labels, act_lens, label_lens, blank)
        acts = acts.contiguous()
        labels, act_lens, label_lens = _contiguous(labels, act_lens, label_lens)
        B, T, U1, V = acts.shape
        dev = acts.device
        f32 = dict(device=dev, dtype=torch.float32)
        costs = torch.empty(B, **f32)
        lp2 = torch.empty(B, T, U1, 2, **f32)
        lse = torch.empty(B, T, U1, **f32)
        alpha, beta = (torch.empty(B, T, U1, device=dev, dtype=torch.int32) for _ in range(2))  # e16m16
        lib = _lib.load()
        with torch.cuda.device(dev):
            _lib.check(lib.rnntb200_loss_dense_fwd(
                _ptr(acts), _DTYPES[acts.dtype], _ptr(labels), _ptr(act_lens), _ptr(label_lens),
                B, T, U1, V, blank, _ptr(costs), _ptr(lp2), _ptr(lse), _ptr(alpha), _ptr(beta),
                _stream()), "rnntb200_loss_dense_fwd")
        ctx.save_for_backward(acts, labels, act_lens, label_lens, lse, alpha, beta)
        ctx.blank = blank
        return costs

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_costs):
        acts, labels, act_lens, label_lens, lse, alpha, beta = ctx.saved_tensors
        B, T, U1, V = acts.shape
        grad_costs = grad_costs.contiguous().to(torch.float32)
        grad = torch.empty_like(acts)
        lib = _lib.load()
        with torch.cuda.device(acts.device):
            _lib.check(lib.rnntb200_loss_dense_bwd(
                _ptr(acts), _DTYPES[acts.dtype], _ptr(labels), _ptr(act_lens), _ptr(label_lens),
                B, T, U1, V, ctx.blank, _ptr(lse), _ptr(alpha), _ptr(beta),
                _ptr(grad_costs), _ptr(grad), _stream()), "rnntb200_loss_dense_bwd")
        return grad, None, None, None, None


class _ConcatGeluRNNT(torch.autograd.Function):
    """costs[B] from the factorised reference joint: logits(t,u) = penc[t] + pdec[u]."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, penc, pdec, labels, act_lens, label_lens, blank, deterministic):
        B, T, V = penc.shape
        U1 = pdec.shape[1]
        _validate((B, T, U1, V), penc.device, labels, act_lens, label_lens, blank)
        penc = penc.contiguous().float()
        pdec = pdec.contiguous().float()
        labels, act_lens, label_lens = _contiguous(labels, act_lens, label_lens)
        f32 = dict(device=penc.device, dtype=torch.float32)
        costs = torch.empty(B, **f32)
        lp2 = torch.empty(B, T, U1, 2, **f32)
        lse = torch.empty(B, T, U1, **f32)
        alpha, beta = (torch.empty(B, T, U1, device=penc.device, dtype=torch.int32) for _ in range(2))
        lib = _lib.load()
        # factor planes (exp of the projections, once per step): written here, read again by the backward
        fac_bytes = lib.rnntb200_joint_cg_factors_bytes(B, T, U1, V)
        factors = torch.empty(fac_bytes, dtype=torch.uint8, device=penc.device)
        with torch.cuda.device(penc.device):
            _lib.check(lib.rnntb200_joint_cg_fwd(
                _ptr(penc), _ptr(pdec), _ptr(labels), _ptr(act_lens), _ptr(label_lens), B, T, U1, V,
                blank, _ptr(costs), _ptr(lp2), _ptr(lse), _ptr(alpha), _ptr(beta),
                _ptr(factors) if fac_bytes else None, fac_bytes, _stream()),
                "rnntb200_joint_cg_fwd")
        ctx.save_for_backward(penc, pdec, labels, act_lens, label_lens, lse, alpha, beta, factors)
        ctx.blank, ctx.deterministic = blank, bool(deterministic)
        return costs

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_costs):
        penc, pdec, labels, act_lens, label_lens, lse, alpha, beta, factors = ctx.saved_tensors
        B, T, V = penc.shape
        U1 = pdec.shape[1]
        grad_costs = grad_costs.contiguous().to(torch.float32)
        fac_bytes = factors.numel()
        d_penc = torch.empty_like(penc)
        d_pdec = torch.empty_like(pdec)
        lib = _lib.load()
        det = int(ctx.deterministic)
        ws_bytes = lib.rnntb200_joint_cg_bwd_workspace_bytes(B, T, U1, V, det)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=penc.device) if ws_bytes else None
        with torch.cuda.device(penc.device):
            _lib.check(lib.rnntb200_joint_cg_bwd(
                _ptr(penc), _ptr(pdec), _ptr(labels), _ptr(act_lens), _ptr(label_lens), B, T, U1, V,
                ctx.blank, _ptr(lse), _ptr(alpha), _ptr(beta), _ptr(grad_costs),
                _ptr(d_penc), _ptr(d_pdec), det, _ptr(ws), ws_bytes,
                _ptr(factors) if fac_bytes else None, fac_bytes, _stream()),
                "rnntb200_joint_cg_bwd")
        return d_penc, d_pdec, None, None, None, None, None


_TICKETS = {}


def _ticket(device):
    """Arrival counter of the fused cost reduction: one int32 that is zero between launches (the sweep re-arms
    it).  Calls must not share a counter unless they are stream-ordered, so every call takes the next of 1024
    slots of a per-device pool that is zeroed once -- no memset per step, nothing keyed on streams, safe under
    CUDA-graph capture (a captured step keeps the slot it was captured with)."""
    pool = _TICKETS.get(device.index)
    if pool is None:
        pool = _TICKETS[device.index] = [torch.zeros(1024, dtype=torch.int32, device=device), 0]
    pool[1] = (pool[1] + 1) % 1024
    return pool[0][pool[1]:pool[1] + 1]


class _ConcatGeluRNNTLoss(torch.autograd.Function):
    """``RNNTLoss(reduction="mean" | "sum")`` of the factorised reference joint as ONE autograd node: the
    sweep's last-arriving utterance adds up the costs (index order: bit-reproducible), the gradient kernel
    takes the one upstream value and scales it itself -- no reduction kernel, no broadcast multiply between
    our kernels (rnntb200_joint_cg_fwd_loss / _bwd_loss).  Returns (loss [1], costs [B] detached)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, penc, pdec, labels, act_lens, label_lens, blank, deterministic, mean):
        B, T, V = penc.shape
        U1 = pdec.shape[1]
        _validate((B, T, U1, V), penc.device, labels, act_lens, label_lens, blank)
        penc = penc.contiguous().float()
        pdec = pdec.contiguous().float()
        labels, act_lens, label_lens = _contiguous(labels, act_lens, label_lens)
        dev = penc.device
        f32 = dict(device=dev, dtype=torch.float32)
        costs, loss = torch.empty(B, **f32), torch.empty(1, **f32)
        lp2 = torch.empty(B, T, U1, 2, **f32)
        lse = torch.empty(B, T, U1, **f32)
        alpha, beta = (torch.empty(B, T, U1, device=dev, dtype=torch.int32) for _ in range(2))
        lib = _lib.load()
        fac_bytes = lib.rnntb200_joint_cg_factors_bytes(B, T, U1, V)
        factors = torch.empty(fac_bytes, dtype=torch.uint8, device=dev)
        scale = 1.0 / B if mean else 1.0
        with torch.cuda.device(dev):
            _lib.check(lib.rnntb200_joint_cg_fwd_loss(
                _ptr(penc), _ptr(pdec), _ptr(labels), _ptr(act_lens), _ptr(label_lens), B, T, U1, V,
                blank, _ptr(costs), _ptr(lp2), _ptr(lse), _ptr(alpha), _ptr(beta),
                _ptr(factors) if fac_bytes else None, fac_bytes, _ptr(loss), _ptr(_ticket(dev)), scale,
                _stream()), "rnntb200_joint_cg_fwd_loss")
        ctx.save_for_backward(penc, pdec, labels, act_lens, label_lens, lse, alpha, beta, factors)
        ctx.blank, ctx.deterministic, ctx.scale = blank, bool(deterministic), scale
        ctx.mark_non_differentiable(costs)
        return loss, costs

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_loss, _grad_costs):
        penc, pdec, labels, act_lens, label_lens, lse, alpha, beta, factors = ctx.saved_tensors
        B, T, V = penc.shape
        U1 = pdec.shape[1]
        grad_loss = grad_loss.reshape(1).contiguous().to(torch.float32)
        fac_bytes = factors.numel()
        d_penc, d_pdec = torch.empty_like(penc), torch.empty_like(pdec)
        lib = _lib.load()
        det = int(ctx.deterministic)
        ws_bytes = lib.rnntb200_joint_cg_bwd_workspace_bytes(B, T, U1, V, det)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=penc.device) if ws_bytes else None
        with torch.cuda.device(penc.device):
            _lib.check(lib.rnntb200_joint_cg_bwd_loss(
                _ptr(penc), _ptr(pdec), _ptr(labels), _ptr(act_lens), _ptr(label_lens), B, T, U1, V,
                ctx.blank, _ptr(lse), _ptr(alpha), _ptr(beta), _ptr(grad_loss), ctx.scale,
                _ptr(d_penc), _ptr(d_pdec), det, _ptr(ws), ws_bytes,
                _ptr(factors) if fac_bytes else None, fac_bytes, _stream()), "rnntb200_joint_cg_bwd_loss")
        return d_penc, d_pdec, None, None, None, None, None, None


class _CgProject(torch.autograd.Function):
    """penc = gelu_tanh(enc) W[:, :He]^T + b,  pdec = gelu_tanh(dec) W[:, He:]^T on the tensor cores at
    fp32 accuracy (rnntb200_joint_cg_project); backward on the tensor cores with the same bf16 hi/lo
    split arithmetic (rnntb200_joint_cg_project_bwd).  enc / dec may be fp32, fp16 or bf16 (the AMP mode:
    the kernels convert on load, d_enc / d_dec come back in the same dtype); weight, bias, penc, pdec and
    the parameter gradients are fp32."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, enc, dec, weight, bias):
        B, T, He = enc.shape
        U1, Hd = dec.shape[1], dec.shape[2]
        V = weight.shape[0]
        lib = _lib.load()
        ws_bytes = lib.rnntb200_joint_cg_project_workspace_bytes(V, He, Hd)
        enc = enc.contiguous()
        dec = dec.contiguous().to(enc.dtype)
        weight, bias = weight.contiguous().float(), bias.contiguous().float()
        penc = torch.empty(B, T, V, device=enc.device, dtype=torch.float32)
        pdec = torch.empty(B, U1, V, device=enc.device, dtype=torch.float32)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=enc.device)
        with torch.cuda.device(enc.device):
            _lib.check(lib.rnntb200_joint_cg_project(
                _ptr(enc), _ptr(dec), _DTYPES[enc.dtype], _ptr(weight), _ptr(bias), B * T, B * U1, He, Hd, V,
                _ptr(penc), _ptr(pdec), _ptr(ws), ws_bytes, _stream()), "rnntb200_joint_cg_project")
        ctx.save_for_backward(enc, dec, weight, ws)  # ws: the weight's bf16 hi/lo split, reused by backward
        return penc, pdec

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, d_penc, d_pdec):
        enc, dec, weight, ws_fwd = ctx.saved_tensors
        He, Hd = enc.shape[-1], dec.shape[-1]
        V = weight.shape[0]
        lib = _lib.load()
        ws_bytes = lib.rnntb200_joint_cg_project_bwd_workspace_bytes(V, He, Hd)
        if ws_bytes > 0:  # tensor-core backward (same hi/lo split arithmetic as the forward)
            d_penc, d_pdec = d_penc.contiguous().float(), d_pdec.contiguous().float()
            d_enc, d_dec = torch.empty_like(enc), torch.empty_like(dec)
            flat = torch.empty(weight.numel() + V, device=enc.device, dtype=torch.float32)  # one buffer: one memset
            d_w, d_b = flat[:weight.numel()].view_as(weight), flat[weight.numel():]
            reuse = ws_fwd.numel() >= ws_bytes
            ws = ws_fwd if reuse else torch.empty(ws_bytes, dtype=torch.uint8, device=enc.device)
            with torch.cuda.device(enc.device):
                _lib.check(lib.rnntb200_joint_cg_project_bwd(
                    _ptr(enc), _ptr(dec), _DTYPES[enc.dtype], _ptr(weight), _ptr(d_penc), _ptr(d_pdec),
                    enc.shape[0] * enc.shape[1], dec.shape[0] * dec.shape[1], He, Hd, V, _ptr(d_enc), _ptr(d_dec),
                    _ptr(d_w), _ptr(d_b), _ptr(ws), ws_bytes, int(reuse), _stream()), "rnntb200_joint_cg_project_bwd")
            return d_enc, d_dec, d_w, d_b
        # widths the tensor-core backward does not tile (not multiples of 128): library GEMMs for this step
        gelu = lambda x: torch.nn.functional.gelu(x, approximate="tanh")
        e32, d32 = enc.float(), dec.float()
        dpe, dpd = d_penc.reshape(-1, V).float(), d_pdec.reshape(-1, V).float()
        d_w = torch.cat((dpe.t() @ gelu(e32).reshape(-1, He), dpd.t() @ gelu(d32).reshape(-1, Hd)), 1)
        d_b = dpe.sum(0)
        d_enc = torch.ops.aten.gelu_backward(d_penc.float() @ weight[:, :He], e32, approximate="tanh").to(enc.dtype)
        d_dec = torch.ops.aten.gelu_backward(d_pdec.float() @ weight[:, He:], d32, approximate="tanh").to(dec.dtype)
        return d_enc, d_dec, d_w, d_b


def project_concat_gelu(enc, dec, weight, bias):
    """(P_enc, P_dec) of the factorised reference joint.  Tensor-core kernel when the shape is
    supported (V <= 80, widths multiples of 64), otherwise library GEMMs."""
    He, Hd, V = enc.size(-1), dec.size(-1), weight.shape[0]
    if enc.is_cuda and enc.dtype in _DTYPES and \
            _lib.load().rnntb200_joint_cg_project_workspace_bytes(V, He, Hd) > 0:
        return _CgProject.apply(enc, dec, weight, bias)
    F = torch.nn.functional
    penc = F.linear(F.gelu(enc.float(), approximate="tanh"), weight[:, :He].float(), bias.float())
    pdec = F.linear(F.gelu(dec.float(), approximate="tanh"), weight[:, He:].float())
    return penc, pdec


_SCALE_VECS = {}


def _scale_vec(device, n, scale):
    """Cached [n] tensor filled with ``scale`` (the backward of mean / sum is one broadcast multiply)."""
    key = (device, n, scale)
    v = _SCALE_VECS.get(key)
    if v is None:
        if len(_SCALE_VECS) > 64:
            _SCALE_VECS.clear()
        v = _SCALE_VECS[key] = torch.full((n,), scale, device=device, dtype=torch.float32)
    return v


class _ReduceCosts(torch.autograd.Function):
    """mean / sum over the utterances as ONE reduction kernel forward and ONE multiply backward
    (eager ``costs.sum() / B`` costs five launches per step across both passes, each as long as its
    launch latency -- 7 % of a cfg-2 step).  The backward hands the fused gradient kernels a
    contiguous fp32 ``[B]`` vector directly."""

    @staticmethod
    def forward(ctx, costs, mean):
        ctx.n, ctx.scale = costs.shape[0], (1.0 / costs.shape[0] if mean else 1.0)
        return costs.mean() if mean else costs.sum()

    @staticmethod
    def backward(ctx, g):
        return g.reshape(1).to(torch.float32) * _scale_vec(g.device, ctx.n, ctx.scale), None


def _reduce(costs, reduction, warp_compat):
    if reduction == "none":
        return costs
    if reduction not in ("sum", "mean"):
        raise ValueError(f"reduction must be 'none', 'mean' or 'sum', got {reduction!r}")
    if costs.is_cuda and costs.dtype == torch.float32 and costs.dim() == 1 and costs.shape[0] > 0:
        out = _ReduceCosts.apply(costs, reduction == "mean")  # warp-transducer: mean over B only
    else:
        out = costs.sum() / costs.shape[0] if reduction == "mean" else costs.sum()
    # warp-transducer returns shape (1,) for mean/sum (model.py:88 torch.cat's them);
    # torchaudio returns a 0-d tensor (model.py:85).
    return out.reshape(1) if warp_compat else out


def rnnt_costs(acts, labels, act_lens, label_lens, blank=0, deterministic=False):
    """Per-utterance costs [B] (differentiable).  ``acts``: dense logits or a JointLogits handle."""
    from .joint import JointLogits  # local import: joint imports loss
    if isinstance(acts, JointLogits):
        return acts.costs(labels, act_lens, label_lens, blank, deterministic)
    if acts.dim() != 4:
        raise RuntimeError("acts must be [B, T, U+1, V]")
    return _DenseRNNT.apply(acts, labels, act_lens, label_lens, int(blank))


def rnnt_loss(acts, labels, act_lens, label_lens, blank=0, reduction="mean", warp_compat=True,
              deterministic=False):
    """Functional form, warp-transducer argument order (north_star)."""
    from .joint import JointLogits  # local import: joint imports loss
    if isinstance(acts, JointLogits) and reduction in ("mean", "sum"):
        return acts.loss(labels, act_lens, label_lens, blank, reduction, warp_compat, deterministic)
    costs = rnnt_costs(acts, labels, act_lens, label_lens, blank, deterministic)
    return _reduce(costs, reduction, warp_compat)


class RNNTLoss(torch.nn.Module):
    """Drop-in for ``warprnnt_pytorch.RNNTLoss`` / ``torchaudio.transforms.RNNTLoss`` as the
    reference uses them (model.py:31,39): ``RNNTLoss(blank, reduction)(acts, labels, act_lens,
    label_lens)``.

    warp_compat=True (default) keeps warp-transducer's ``(1,)`` result shape for mean/sum, which
    ``validation_epoch_end`` relies on (model.py:83-88); False gives torchaudio's 0-d tensor.
    """

    def __init__(self, blank: int = 0, reduction: str = "mean", warp_compat: bool = True,
                 deterministic: bool = False, check_lengths: bool = False):
        super().__init__()
        if reduction not in ("none", "mean", "sum"):
            raise ValueError(f"reduction must be 'none', 'mean' or 'sum', got {reduction!r}")
        self.blank = int(blank)
        self.reduction = reduction
        self.warp_compat = warp_compat
        self.deterministic = deterministic
        self.check_lengths = check_lengths  # debug only: costs a device sync (lengths AND label range)

    def forward(self, acts, labels, act_lens, label_lens):
        if self.check_lengths:
            shape = acts.shape
            check_lengths(act_lens, label_lens, shape[1], shape[2], labels, shape[3])
        return rnnt_loss(acts, labels, act_lens, label_lens, self.blank, self.reduction,
                         self.warp_compat, self.deterministic)
