"""CPU: a Python model of the BOOKKEEPING of the self-contained lattice sweep (csrc/lattice.cu, "tp").

One warp does everything for its 32 label positions; its global accesses are row-wise through two
lane-private FIFOs whose slot index is the step index.  What can go wrong there is index arithmetic: which
FIFO slot holds which lattice row of which lane, how far the loader may run ahead before it overwrites a row
a late lane still needs, when a row of the output plane is complete, which edge-ring slot a neighbouring
warp's value sits in, which per-step slot of the next band's array a boundary value goes to.  This model
replays exactly that (same constants, same index expressions, every read checked against a tag that says what
the slot holds) with plain log-domain arithmetic and compares the planes with the fp64 oracle.  The arithmetic
of the chain has its own model (tests/test_chain_arithmetic.py); the CUDA code itself is checked on the GPU.
"""
import numpy as np
import pytest

from oracle import np_oracle

KB, AHEAD, RAW, OUT, EDGE, BAND_SKEW = 8, 2, 64, 64, 32, 16  # kTpKB, kTpAhead, kTpRaw, kTpOut, kTpEdge, kTpBandSkew
NEG = -np.inf


def log_softmax(x):
    m = x.max(-1, keepdims=True)
    return x - m - np.log(np.exp(x - m).sum(-1, keepdims=True))


class Warp:
    def __init__(self, wg, band, nw, direction, lp2, Tb, Ub):
        self.wg, self.band, self.w, self.nw, self.dir = wg, band, wg - band * nw, nw, direction
        self.lp2, self.Tb, self.Ub = lp2, Tb, Ub
        self.lag = wg * KB + band * (BAND_SKEW - KB)
        self.base = self.lag + 32 * wg
        self.j = wg * 32 + np.arange(32)
        self.on = self.j <= Ub
        self.raw = np.zeros((RAW, 32, 2))
        self.raw_tag = np.full((RAW, 32), -10**9)
        self.out = np.zeros((OUT, 32))
        self.out_tag = np.full((OUT, 32), -10**9)
        self.r_ld = -self.base
        self.row_st = -self.base - 31
        self.slot0 = (-self.base) % RAW
        self.own = np.where(self.j == 0, 0.0, NEG) if direction == 0 else np.full(32, NEG)
        self.share = np.full(32, NEG)
        self.prev = np.where(self.j == 0, 0.0, NEG)  # beta: val(-1) = 1 on lane j = 0
        for _ in range(AHEAD + 1):
            self.load_block()

    def cell(self, lane, tau):
        if not self.on[lane] or not (0 <= tau < self.Tb):
            return None
        return (tau, self.j[lane]) if self.dir == 0 else (self.Tb - 1 - tau, self.Ub - self.j[lane])

    def load_block(self):
        for k in range(KB):
            r = self.r_ld + k
            rc = min(max(r, 0), self.Tb - 1)  # clamped, not skipped
            for lane in range(32):
                if self.on[lane]:
                    slot = (r + lane) % RAW
                    c = self.cell(lane, rc)
                    self.raw[slot, lane] = self.lp2[c]
                    self.raw_tag[slot, lane] = r
        self.r_ld += KB

    def factors(self, blk):
        """log factors of block blk, read from the step-indexed slots (lanes without a cell: whatever is there)."""
        slot = (blk * KB - self.base) % RAW
        assert slot == self.slot0, "the chain side addresses both FIFOs with one register that advances by KB per block"
        f = np.zeros((KB, 32, 2))
        for k in range(KB):
            for lane in range(32):
                tau = blk * KB + k - self.lag - self.j[lane]
                if self.cell(lane, tau) is not None:
                    assert self.raw_tag[slot + k, lane] == tau, "FIFO slot overwritten or row not loaded yet"
                    f[k, lane] = self.raw[slot + k, lane]
        return f


def sweep(lp2, Tb, Ub, nw, direction):
    n_on = (Ub + 32) // 32
    n_bands = (n_on + nw - 1) // nw
    max_lag = (n_on - 1) * KB + ((n_on - 1) // nw) * (BAND_SKEW - KB)
    n_blocks = (Tb + Ub + max_lag + KB - 1) // KB
    plane = np.full((Tb, Ub + 1), np.nan)
    warps = [Warp(b * nw + w, b, nw, direction, lp2, Tb, Ub) for b in range(n_bands) for w in range(nw)]
    edge = {b: (np.full((EDGE, 8), NEG), np.full((EDGE, 8), -10**9)) for b in range(n_bands)}
    xedge = {b: {} for b in range(n_bands)}  # per-step slots, written once by the previous band
    for blk in range(n_blocks):
        for W in warps:  # within a block the warps run in any order: a barrier separates the blocks
            W.load_block()  # rows of block blk + AHEAD + 1
            f = W.factors(blk)
            ring, tag = edge[W.band]
            er = (blk * KB - W.lag) % EDGE
            ev = np.full(KB, NEG)
            if W.w > 0:
                for k in range(KB):
                    q = blk * KB + k - W.lag - 1
                    if q >= 0:
                        assert tag[er + k, W.w - 1] == q, "edge-ring slot overwritten or not yet written"
                        ev[k] = ring[er + k, W.w - 1]
            elif W.band > 0:
                for k in range(KB):
                    q = blk * KB + k - W.lag - 1
                    if q >= 0:
                        assert q in xedge[W.band], "band boundary value not forwarded yet (the receiver would spin)"
                        ev[k] = xedge[W.band][q]
            new_edges = []
            for k in range(KB):
                inn = np.concatenate(([ev[k]], (W.share if direction == 0 else W.prev)[:-1]))
                if direction == 0:
                    v = np.logaddexp(W.own, inn)
                    W.own = v + f[k, :, 0]
                    W.share = v + f[k, :, 1]
                    shared = W.share[31]
                else:
                    first = blk == 0 and k == 0
                    inn_term = np.full(32, NEG) if first else inn + f[k, :, 1]
                    inn_term[0] = ev[k] + f[k, 0, 1]
                    v = np.logaddexp(W.prev + f[k, :, 0], inn_term)
                    W.prev = v
                    shared = v[31]
                slot = (W.slot0 + k) % OUT
                tau0 = blk * KB + k - W.lag
                for lane in range(32):
                    W.out[slot, lane] = v[lane]
                    W.out_tag[slot, lane] = tau0 - W.j[lane]
                p = blk * KB + k - W.lag  # this warp's local step: stored at edge slot p + 1
                es = (p + 1) % EDGE
                ring[es, W.w], tag[es, W.w] = shared, p
                new_edges.append((p, shared))
            if W.w == nw - 1 and W.band + 1 < n_bands:  # forwarded once per block, from the edge ring
                ew = (blk * KB - W.lag) % EDGE
                for lane in range(KB):
                    q = blk * KB + lane - W.lag
                    if q >= 0:
                        s = ew + lane + 1 if lane + 1 < KB else (ew + KB) % EDGE
                        assert tag[s, W.w] == q
                        assert q not in xedge[W.band + 1], "a per-step slot is written exactly once"
                        xedge[W.band + 1][q] = ring[s, W.w]
            for k in range(KB):  # rows lane 31 has passed
                r = W.row_st + k
                for lane in range(32):
                    c = W.cell(lane, r)
                    if c is not None:
                        assert W.out_tag[(r + lane) % OUT, lane] == r, "output row stored before it was complete"
                        plane[c] = W.out[(r + lane) % OUT, lane]
            W.row_st += KB
            W.slot0 = (W.slot0 + KB) % RAW
    for W in warps:
        for r in range(W.row_st, Tb):
            for lane in range(32):
                c = W.cell(lane, r)
                if c is not None:
                    assert W.out_tag[(r + lane) % OUT, lane] == r
                    plane[c] = W.out[(r + lane) % OUT, lane]
    return plane


@pytest.mark.parametrize("T,U,nw", [(40, 20, 1), (37, 70, 3), (90, 95, 3), (21, 130, 2), (70, 150, 2), (33, 200, 4),
                                    (5, 40, 2), (150, 3, 1), (64, 127, 4)])
def test_tp_bookkeeping_reproduces_the_oracle(T, U, nw):
    rng = np.random.default_rng(T * 1000 + U)
    V = 5
    logits = rng.normal(size=(1, T, U + 1, V)) * 2.0
    labels = rng.integers(1, V, size=(1, U)).astype(np.int32)
    ref = np_oracle.rnnt_loss_np(logits, labels, [T], [U], 0)
    lp = log_softmax(logits[0])
    lab = np.concatenate([labels[0], [0]]).astype(np.int64)
    lp2 = np.stack([lp[..., 0], np.take_along_axis(lp, lab[None, :, None].repeat(T, 0), 2)[..., 0]], -1)
    n_on = (U + 32) // 32
    warps_per_cta = nw if n_on > 4 else n_on  # up to four warps: one CTA; beyond: bands of nw warps in a cluster
    alpha = sweep(lp2, T, U, warps_per_cta, 0)
    beta = sweep(lp2, T, U, warps_per_cta, 1)
    assert not np.isnan(alpha).any() and not np.isnan(beta).any(), "a lattice cell was never stored"
    np.testing.assert_allclose(alpha, ref["alphas"][0], rtol=0, atol=1e-9)
    np.testing.assert_allclose(beta, ref["betas"][0], rtol=0, atol=1e-9)
    np.testing.assert_allclose(-beta[0, 0], ref["costs"][0], rtol=1e-12)
