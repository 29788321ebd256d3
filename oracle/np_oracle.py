"""fp64 numpy restatement of the RNN-T loss maths (SURVEY.md 8(a) "Exact maths to
implement") -- an independent second opinion next to ``warp_cpu.c`` for SMALL cases
(pure-Python loops over the lattice).  TEST INFRASTRUCTURE ONLY.

Gradient here is written in the closed form w.r.t. the LOGITS (the form torchaudio's
and warp-transducer's GPU kernels use), i.e. not via log-softmax backward, so the two
oracles check each other's algebra.
"""
from __future__ import annotations

import numpy as np


def _log_softmax(x):
    m = x.max(axis=-1, keepdims=True)
    return x - m - np.log(np.exp(x - m).sum(axis=-1, keepdims=True))


def rnnt_loss_np(logits, labels, act_lens, label_lens, blank=0):
    x = np.asarray(logits, dtype=np.float64)
    B, Tm, U1m, V = x.shape
    costs = np.zeros(B)
    grads = np.zeros_like(x)
    alphas = np.zeros((B, Tm, U1m))
    betas = np.zeros((B, Tm, U1m))
    for b in range(B):
        T, U = int(act_lens[b]), int(label_lens[b])
        U1 = U + 1
        y = np.asarray(labels[b][:U], dtype=np.int64)
        lp = _log_softmax(x[b, :T, :U1])
        a = np.full((T, U1), -np.inf)
        a[0, 0] = 0.0
        for t in range(T):
            for u in range(U1):
                if t == 0 and u == 0:
                    continue
                ne = a[t - 1, u] + lp[t - 1, u, blank] if t > 0 else -np.inf
                em = a[t, u - 1] + lp[t, u - 1, y[u - 1]] if u > 0 else -np.inf
                a[t, u] = np.logaddexp(ne, em)
        be = np.full((T, U1), -np.inf)
        be[T - 1, U] = lp[T - 1, U, blank]
        for t in range(T - 1, -1, -1):
            for u in range(U, -1, -1):
                if t == T - 1 and u == U:
                    continue
                ne = be[t + 1, u] + lp[t, u, blank] if t < T - 1 else -np.inf
                em = be[t, u + 1] + lp[t, u, y[u]] if u < U else -np.inf
                be[t, u] = np.logaddexp(ne, em)
        ll = be[0, 0]
        costs[b] = -ll
        g = np.exp(a[:, :, None] + be[:, :, None] + lp - ll)
        for t in range(T):
            for u in range(U1):
                if t < T - 1:
                    g[t, u, blank] -= np.exp(a[t, u] + lp[t, u, blank] + be[t + 1, u] - ll)
                elif u == U:
                    g[t, u, blank] -= np.exp(a[t, u] + lp[t, u, blank] - ll)
                if u < U:
                    g[t, u, y[u]] -= np.exp(a[t, u] + lp[t, u, y[u]] + be[t, u + 1] - ll)
        grads[b, :T, :U1] = g
        alphas[b, :T, :U1] = a
        betas[b, :T, :U1] = be
    return dict(costs=costs, grads=grads, alphas=alphas, betas=betas)
