"""CPU: numpy model of the ARITHMETIC of the lattice sweep's chain warps (csrc/lattice.cu, "decoupled
(mantissa | exponent) recursion").

tests/test_sweep_schedule.py replays the kernel's bookkeeping with log-domain arithmetic; this file
replays its arithmetic with trivial bookkeeping: values are (fp32 mantissa, int32 exponent) pairs, the
exponents follow an integer max-plus recurrence, alignment scales are 2^-(gap) clipped to 0 at 127, lane
0's shuffled term is switched off by an exponent bias, the value crossing a warp boundary enters as a
third term, mantissas are renormalised once per block of KB steps.  Same constants and the same
expressions as the CUDA code (int32 overflow is checked explicitly), compared with the fp64 oracle.
"""
import numpy as np
import pytest

from oracle import np_oracle

KB = 8
K_ZERO = -(1 << 29)
K_NOTERM = K_ZERO
LOG2E = np.float32(1.4426950408889634)
f32 = np.float32


def log_softmax(x):
    m = x.max(-1, keepdims=True)
    return x - m - np.log(np.exp(x - m).sum(-1, keepdims=True))


def me_from_log(lp):
    x = np.maximum(f32(lp) * LOG2E, f32(-16000.0)).astype(np.float32)
    t = np.rint(x).astype(np.float32)
    return np.exp2((x - t).astype(np.float32)).astype(np.float32), t.astype(np.int64)


def pow2_neg(d):
    assert (d >= 0).all(), "alignment gap must be non-negative"
    k = np.minimum(d, 127)
    return np.where(k >= 127, f32(0), np.exp2(-k.astype(np.float64))).astype(np.float32)


def i32(x):
    x = np.asarray(x, dtype=np.int64)
    assert (np.abs(x) < (1 << 31)).all(), "int32 overflow in the exponent recurrence"
    return x


def shfl_up(x):
    return np.concatenate((x[:1], x[:-1]))  # lane 0 gets its own value back, like __shfl_up_sync


def sweep(lp2, Tb, Ub, direction):
    """Planes (mantissa, exponent) -> natural-log values, one utterance, ceil((Ub+1)/32) chain warps."""
    n_warps = (Ub + 32) // 32
    n_blocks = (Tb + Ub + (n_warps - 1) * KB + KB - 1) // KB
    S = n_blocks * KB
    plane = np.full((Tb, Ub + 1), np.nan)
    lanes = np.arange(32)
    edge_out = [dict() for _ in range(n_warps)]  # warp w: step-local diagonal index -> (m, E) of lane 31's share

    def cell(j, tau):
        if j > Ub or not (0 <= tau < Tb):
            return None
        return (tau, j) if direction == 0 else (Tb - 1 - tau, Ub - j)

    for w in range(n_warps):  # warp w-1 is always a whole block ahead: run the warps one after the other
        lag = w * KB
        j = w * 32 + lanes
        m = np.ones(32, dtype=np.float32)
        E = np.where(j == 0, 0, K_ZERO).astype(np.int64)
        pbm_prev = np.ones(32, dtype=np.float32)
        plm_prev = np.ones(32, dtype=np.float32)
        pbe_prev = np.zeros(32, dtype=np.int64)
        ple_prev = np.full(32, K_NOTERM if direction == 0 else 0, dtype=np.int64)
        in_bias = np.where(lanes == 0, K_NOTERM, 0).astype(np.int64)
        for s in range(S):
            k = s % KB
            tau = s - lag - j
            fm = np.ones((32, 2), dtype=np.float32)
            fe = np.zeros((32, 2), dtype=np.int64)
            for lane in range(32):
                c = cell(j[lane], tau[lane])
                if c is not None:
                    for comp in range(2):
                        fm[lane, comp], fe[lane, comp] = me_from_log(lp2[c][comp])
            evm = np.ones(32, dtype=np.float32)
            evE = np.full(32, K_NOTERM, dtype=np.int64)
            if w > 0:
                q = s - lag - 1
                evm[0], evE[0] = edge_out[w - 1].get(q, (f32(1), K_ZERO))
            if direction == 0:
                pbe, ple, pbm, plm_in, plm_edge = pbe_prev, ple_prev, pbm_prev, f32(1), f32(1)
                oE = i32(E + pbe)
                iE = i32(shfl_up(i32(E + ple)) + in_bias)
                eE = i32(evE)
                shm = (m * plm_prev).astype(np.float32)
            else:
                pbe, ple, pbm, plm_in, plm_edge = fe[:, 0], fe[:, 1], fm[:, 0], fm[:, 1], fm[:, 1]
                oE = i32(E + pbe)
                iE = i32(shfl_up(E) + (K_NOTERM if s == 0 else in_bias) + ple)  # beta's seed must not reach lane 1
                eE = i32(evE + ple)
                shm = m
            En = np.maximum(np.maximum(oE, iE), eE)
            c_own = (pow2_neg(i32(En - oE)) * pbm).astype(np.float32)
            c_in = (pow2_neg(i32(En - iE)) * plm_in).astype(np.float32)
            c_edge = (pow2_neg(i32(En - eE)) * plm_edge).astype(np.float32)
            in_m = shfl_up(shm)
            m = (in_m * c_in + (evm * c_edge + m * c_own)).astype(np.float32)
            E = En
            assert np.isfinite(m).all() and (m > 0).all()
            assert (m < f32(2.0) ** 40).all() and (m > f32(2.0) ** -40).all(), "mantissa drifted out of range"
            if k == KB - 1:
                mant, ex = np.frexp(m)  # m = mant * 2^ex, mant in [0.5, 1)
                m, E = (mant * 2).astype(np.float32), i32(E + ex - 1)
            for lane in range(32):
                c = cell(j[lane], tau[lane])
                if c is not None:
                    plane[c] = (float(E[lane]) + np.log2(float(m[lane]))) * np.log(2.0)
            pbm_prev, plm_prev, pbe_prev, ple_prev = fm[:, 0], fm[:, 1], fe[:, 0], fe[:, 1]
            share_m = (m * fm[:, 1]).astype(np.float32) if direction == 0 else m
            share_E = i32(E + fe[:, 1]) if direction == 0 else E
            edge_out[w][s - lag] = (share_m[31], int(share_E[31]))
    return plane


@pytest.mark.parametrize("T,U,scale", [(40, 20, 2.0), (37, 70, 2.0), (25, 95, 1.0), (150, 3, 3.0), (1, 5, 1.0),
                                       (30, 0, 2.0), (60, 40, 60.0)])
def test_decoupled_recursion_matches_the_oracle(T, U, scale):
    rng = np.random.default_rng(T * 1000 + U)
    V = 5
    logits = rng.normal(size=(1, T, U + 1, V)) * scale  # scale 60: per-step factors down to 2^-400
    labels = rng.integers(1, V, size=(1, max(U, 1))).astype(np.int32)[:, :U]
    ref = np_oracle.rnnt_loss_np(logits, labels, [T], [U], 0)
    lp = log_softmax(logits[0])
    lab = np.concatenate([labels[0], [0]]).astype(np.int64)
    lp2 = np.stack([lp[..., 0], np.take_along_axis(lp, lab[None, :, None].repeat(T, 0), 2)[..., 0]], -1)
    for direction, key in ((0, "alphas"), (1, "betas")):
        got = sweep(lp2, T, U, direction)
        assert not np.isnan(got).any(), "a lattice cell was never produced"
        # fp32 mantissas: ~1e-7 relative per step -> |d ln| <= ~(T+U) * 2e-7; fp32 log-probs in: |lp| * 6e-8 per factor
        tol = (T + U) * 4e-7 * max(1.0, scale * 4)
        np.testing.assert_allclose(got, ref[key][0], rtol=0, atol=tol)
