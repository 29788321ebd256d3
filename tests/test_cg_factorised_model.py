"""CPU: a numpy model of the FACTORISED concat-GELU cell kernels (csrc/joint_cg_mm.cu) and of their wide
(128-column chunk) decomposition, against the fp64 numpy oracle.

logits(t,u,v) = P_enc[t,v] + P_dec[u,v], so with A = 2^((P_enc - rowmax) log2 e), B likewise,
    partition  S = A B^T                                   (the vocabulary is the K dimension)
    dP_enc     = A .* (C B)   - X,   X[t,v] = sum_u cl(t,u) [v = y_u] + cb(t,u) [v = blank]
    dP_dec     = B .* (C^T A) - column sums of cb at the blank column, of cl at the label column
with C = grad_cost * occupancy / S and occupancy = (cb + cl) / grad_cost by the beta recursion.  The wide
kernels evaluate the same products in 128-column chunks: the partition ACCUMULATES over the chunks, the
gradient is computed per chunk independently with GLOBAL column indices in the one-hot corrections.  This
model replays exactly that decomposition (same chunk width, partial last chunk, padded columns = 0) in fp64;
what it pins on the CPU is the algebra and the index bookkeeping, the CUDA code itself is checked on the GPU
(tests/test_gpu_joint_cg_wide.py).
"""
import numpy as np
import pytest

from oracle import np_oracle

WC = 128  # kWC: vocabulary columns per chunk


def factor_planes(P):
    """cg_factor_rows*: E = exp(P - rowmax) with the row padded to a multiple of 8 by zeros, and the row maxima."""
    V = P.shape[-1]
    Vk = (V + 7) & ~7
    m = P.max(-1)
    E = np.zeros(P.shape[:-1] + (Vk,))
    E[..., :V] = np.exp(P - m[..., None])
    return E, m, Vk


def factorised_step(penc, pdec, y, blank, grad_cost=1.0):
    """One utterance (T frames, U labels): costs-independent pieces of the forward + the whole backward of the
    cell kernels, column chunk by column chunk.  alpha / beta come from the oracle recursion on the model's own
    log-probabilities (the sweep has its own tests)."""
    T, V = penc.shape
    U1 = pdec.shape[0]
    U = U1 - 1
    A, mA, Vk = factor_planes(penc)
    Bm, mB, _ = factor_planes(pdec)
    n_vc = (Vk + WC - 1) // WC
    # forward: S accumulates over the column chunks (K loop)
    S = np.zeros((T, U1))
    for vc in range(n_vc):
        v0, w = vc * WC, min(WC, Vk - vc * WC)
        assert w % 8 == 0 and w > 0
        S += A[:, v0:v0 + w] @ Bm[:, v0:v0 + w].T
    lse = mA[:, None] + mB[None, :] + np.log(S)
    lp_blank = penc[:, blank][:, None] + pdec[:, blank][None, :] - lse
    lp_label = np.zeros((T, U1))
    lp_label[:, :U] = penc[:, y] + pdec[np.arange(U), y][None, :] - lse[:, :U]
    # alpha / beta of these log-probs (plain log-domain recursion)
    al = np.full((T, U1), -np.inf)
    al[0, 0] = 0.0
    for t in range(T):
        for u in range(U1):
            if t == 0 and u == 0:
                continue
            ne = al[t - 1, u] + lp_blank[t - 1, u] if t > 0 else -np.inf
            em = al[t, u - 1] + lp_label[t, u - 1] if u > 0 else -np.inf
            al[t, u] = np.logaddexp(ne, em)
    be = np.full((T, U1), -np.inf)
    be[T - 1, U] = lp_blank[T - 1, U]
    for t in range(T - 1, -1, -1):
        for u in range(U, -1, -1):
            if t == T - 1 and u == U:
                continue
            ne = be[t + 1, u] + lp_blank[t, u] if t < T - 1 else -np.inf
            em = be[t, u + 1] + lp_label[t, u] if u < U else -np.inf
            be[t, u] = np.logaddexp(ne, em)
    ll = be[0, 0]
    # per-cell scalars of the gradient kernel: cb, cl, C (beta(t,u) itself is not read)
    cb = np.zeros((T, U1))
    cl = np.zeros((T, U1))
    cb[:T - 1] = grad_cost * np.exp(al[:T - 1] + be[1:] + lp_blank[:T - 1] - ll)
    cb[T - 1, U] = grad_cost * np.exp(al[T - 1, U] + lp_blank[T - 1, U] - ll)
    cl[:, :U] = grad_cost * np.exp(al[:, :U] + be[:, 1:] + lp_label[:, :U] - ll)
    C = (cb + cl) * np.exp(mA[:, None] + mB[None, :] - lse)  # (cb + cl) / S
    # backward, one column chunk at a time (blockIdx.z), global indices in the corrections
    d_penc = np.zeros((T, V))
    d_pdec = np.zeros((U1, V))
    for vc in range(n_vc):
        v_off, Vc = vc * WC, min(WC, Vk - vc * WC)
        Ac, Bc = A[:, v_off:v_off + Vc], Bm[:, v_off:v_off + Vc]
        E = C @ Bc        # [T, Vc]
        D = C.T @ Ac      # [U1, Vc]
        X = np.zeros((T, Vc))
        for col in range(Vc):
            v = v_off + col
            for u in range(U):
                if y[u] == v:
                    X[:, col] += cl[:, u]
            if v == blank:
                X[:, col] += cb.sum(1)
        for col in range(Vc):
            v = v_off + col
            if v >= V:
                assert np.all(Ac[:, col] == 0) and np.all(Bc[:, col] == 0)  # pad columns contribute nothing
                continue
            d_penc[:, v] = Ac[:, col] * E[:, col] - X[:, col]
            g = Bc[:, col] * D[:, col]
            if v == blank:
                g = g - cb.sum(0)
            for u in range(U):
                if y[u] == v:
                    g[u] -= cl[:, u].sum()
            d_pdec[:, v] = g
    return -ll, d_penc, d_pdec


@pytest.mark.parametrize("T,U,V,blank", [(9, 5, 73, 0), (7, 4, 129, 0), (6, 6, 136, 130), (5, 3, 257, 128), (4, 7, 300, 299)])
def test_factorised_chunked_model_matches_the_oracle(T, U, V, blank):
    rng = np.random.default_rng(V * 100 + T)
    penc = rng.normal(size=(T, V)) * 2.0
    pdec = rng.normal(size=(U + 1, V)) * 2.0
    y = rng.integers(0, V - 1, size=U)
    y = y + (y >= blank)  # any column but the blank
    y[0] = V - 1 if blank != V - 1 else V - 2  # a label in the last (partial) chunk
    logits = penc[None, :, None, :] + pdec[None, None, :, :]
    ref = np_oracle.rnnt_loss_np(logits, y[None, :], [T], [U], blank)
    cost, d_penc, d_pdec = factorised_step(penc, pdec, y, blank)
    np.testing.assert_allclose(cost, ref["costs"][0], rtol=1e-12)
    np.testing.assert_allclose(d_penc, ref["grads"][0].sum(1), atol=1e-12)  # dP_enc[t,v] = sum_u dlogits[t,u,v]
    np.testing.assert_allclose(d_pdec, ref["grads"][0].sum(0), atol=1e-12)  # dP_dec[u,v] = sum_t dlogits[t,u,v]
