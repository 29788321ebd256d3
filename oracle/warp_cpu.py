"""ctypes front-end of ``oracle/librnnt_oracle.so`` (C restatement of warp-transducer's
CPU RNN-T loss; see ``warp_cpu.c``).  TEST INFRASTRUCTURE ONLY.

``rnnt_loss_cpu`` mirrors the call the reference issues at ``model.py:57,74``
(``loss(acts, labels, act_lens, label_lens)`` with ``blank``/``reduction`` from
``model.py:39``) on numpy arrays, returning per-utterance costs and the gradient of
``sum(costs)`` w.r.t. the logits.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "librnnt_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C oracle in place (gcc + OpenMP).  Idempotent."""
    src = os.path.join(_HERE, "warp_cpu.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B" if force else "-s"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        lib = ctypes.CDLL(_SO)
        for suffix, ctype in (("f32", ctypes.c_float), ("f64", ctypes.c_double)):
            fn = getattr(lib, f"rnnt_oracle_cost_and_grad_{suffix}")
            p = ctypes.POINTER(ctype)
            ip = ctypes.POINTER(ctypes.c_int)
            fn.argtypes = [p, ip, ip, ip] + [ctypes.c_int] * 5 + [p, p, p, p, ctypes.c_int]
            fn.restype = ctypes.c_int
        lib.rnnt_oracle_max_threads.restype = ctypes.c_int
        _lib = lib
    return _lib


def max_threads() -> int:
    return int(_load().rnnt_oracle_max_threads())


def rnnt_loss_cpu(logits, labels, act_lens, label_lens, blank=0, want_grad=True,
                  want_alpha_beta=False, num_threads=0, dtype=np.float32):
    """Per-utterance costs (and d sum(costs)/d logits) of the RNN-T loss on the CPU.

    logits [B,T,U+1,V] float; labels [B,U] int32; act_lens/label_lens [B] int32.
    Returns dict(costs[B], grads[B,T,U+1,V] | None, alphas/betas [B,T,U+1] | None).
    """
    lib = _load()
    dtype = np.dtype(dtype)
    suffix, ctype = ("f32", ctypes.c_float) if dtype == np.float32 else ("f64", ctypes.c_double)
    x = np.ascontiguousarray(logits, dtype=dtype)
    if x.ndim != 4:
        raise ValueError("logits must be [B,T,U+1,V]")
    B, T, U1, V = x.shape
    lab = np.ascontiguousarray(labels, dtype=np.int32).reshape(B, max(U1 - 1, 0))
    al = np.ascontiguousarray(act_lens, dtype=np.int32)
    ll = np.ascontiguousarray(label_lens, dtype=np.int32)
    if lab.size == 0:
        lab = np.zeros((B, 1), dtype=np.int32)  # never dereferenced when U == 0
    costs = np.zeros(B, dtype=dtype)
    grads = np.zeros_like(x) if want_grad else None
    alphas = np.zeros((B, T, U1), dtype=dtype) if want_alpha_beta else None
    betas = np.zeros((B, T, U1), dtype=dtype) if want_alpha_beta else None
    p = ctypes.POINTER(ctype)
    ip = ctypes.POINTER(ctypes.c_int)

    def ptr(a, t):
        return a.ctypes.data_as(t) if a is not None else None

    nt = int(num_threads) if num_threads else min(max_threads(), max(B, 1))
    st = getattr(lib, f"rnnt_oracle_cost_and_grad_{suffix}")(
        ptr(x, p), ptr(lab, ip), ptr(al, ip), ptr(ll, ip), B, T, U1, V, int(blank),
        ptr(costs, p), ptr(grads, p), ptr(alphas, p), ptr(betas, p), nt)
    if st != 0:
        raise ValueError(f"rnnt oracle: invalid arguments (status {st})")
    return dict(costs=costs, grads=grads, alphas=alphas, betas=betas, threads=nt)


def reduce_costs(costs, reduction="mean"):
    """warp-transducer's reductions: ``mean`` divides by B only."""
    if reduction == "none":
        return costs
    if reduction == "sum":
        return costs.sum(keepdims=True)
    if reduction == "mean":
        return costs.sum(keepdims=True) / costs.shape[0]
    raise ValueError(reduction)
