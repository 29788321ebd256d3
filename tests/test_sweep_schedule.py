"""CPU: a Python model of the warp-specialised lattice sweep's SCHEDULE (csrc/lattice.cu, second half).

The CUDA kernel splits one sweep over chain warps, helper warps and -- for long label sequences -- the
CTAs of a cluster.  What can go wrong there is bookkeeping, not arithmetic: which step a lane's cell
belongs to, which ring slot a neighbour's value is read from, when a shared-memory window row may be
overwritten, which rows are complete when they are stored.  This model replays exactly that
bookkeeping (same constants, same index expressions, every ring with the maximal run-ahead / lag the
barriers allow) with plain log-domain arithmetic, asserts every slot read finds what the index math
promises, and compares the planes it produces with the fp64 oracle.  It guards the design; the CUDA
code itself is checked on the GPU (tests/test_gpu_loss.py::test_sweep_warp_boundaries).
"""
import numpy as np
import pytest

from oracle import np_oracle

KB, RW, WIN, EDGE, SKEW, STAGES = 8, 128, 64, 64, 32, 3  # kUnroll-block, loader window, consumer window,
NEG = -np.inf                                             # edge ring, band skew, ring stages (lattice.cu)
K_RUN = RW // KB - 5                                       # blocks the loader may run ahead of the chain


def log_softmax(x):
    m = x.max(-1, keepdims=True)
    return x - m - np.log(np.exp(x - m).sum(-1, keepdims=True))


class Band:
    """One CTA: nw chain warps with their rings and windows."""

    def __init__(self, band, n_bands, nw, direction, lp2, Tb, Ub):
        self.band, self.n_bands, self.nw, self.dir, self.lp2, self.Tb, self.Ub = band, n_bands, nw, direction, lp2, Tb, Ub
        self.xlag = band * (SKEW - KB)
        self.own = np.full((nw, 32), NEG)
        self.share = np.full((nw, 32), NEG)
        self.edge = np.full((EDGE, nw), NEG)
        self.edge_tag = np.full((EDGE, nw), -10**9)
        self.val = np.zeros((nw, STAGES, KB, 32))
        self.raw = np.zeros((nw, RW, 32, 2))
        self.raw_row = np.full((nw, RW, 32), -10**9)      # which lattice row a window slot holds
        self.loaded_blocks = [0] * nw
        self.out = np.zeros((nw, WIN, 32))
        self.out_row = np.full((nw, WIN, 32), -10**9)
        self.consumed = [0] * nw                          # value blocks the consumer has drained
        self.row_st = [0] * nw
        self.xedge = {}                                   # filled by the previous band's consumer
        self.xdone = 0
        for w in range(nw):
            wg = band * nw + w
            if wg == 0:
                self.own[w, 0] = 0.0
            self.row_st[w] = -(self.lag(w) + 32 * wg) - 31

    def lag(self, w):
        return (self.band * self.nw + w) * KB + self.xlag

    def cell(self, w, lane, tau):
        j = (self.band * self.nw + w) * 32 + lane
        if j > self.Ub or not (0 <= tau < self.Tb):
            return None
        return (tau, j) if self.dir == 0 else (self.Tb - 1 - tau, self.Ub - j)

    # ---- loader: rows of block pb of warp w -> window (allowed while chain_done > pb - K_RUN)
    def load_ahead(self, w, chain_done, n_blocks):
        base = self.lag(w) + 32 * (self.band * self.nw + w)
        while self.loaded_blocks[w] < n_blocks and (self.loaded_blocks[w] < K_RUN or chain_done > self.loaded_blocks[w] - K_RUN):
            pb = self.loaded_blocks[w]
            for k in range(KB):
                r = pb * KB + k - base
                for lane in range(32):
                    c = self.cell(w, lane, r)
                    if c is not None:
                        self.raw[w, r & (RW - 1), lane] = self.lp2[c]
                        self.raw_row[w, r & (RW - 1), lane] = r
            self.loaded_blocks[w] += 1

    # ---- chain: block blk of warp w
    def chain_block(self, w, blk, prev_band):
        lag = self.lag(w)
        j0 = (self.band * self.nw + w) * 32
        for k in range(KB):
            s = blk * KB + k
            inn = np.concatenate(([NEG], self.share[w, :-1]))           # shfl_up by one lane
            q = s - lag - 1                                              # slot of diagonal d-1
            if w > 0:
                if q >= 0:
                    assert self.edge_tag[q & (EDGE - 1), w - 1] == q, "edge ring slot overwritten or not yet written"
                    inn[0] = self.edge[q & (EDGE - 1), w - 1]
            elif self.band > 0:
                if q >= 0:
                    assert blk < SKEW // KB or self.xdone > blk - SKEW // KB, "previous band has not forwarded this block"
                    assert q in self.xedge, "boundary value not forwarded"
                    inn[0] = self.xedge[q]
            fac = np.zeros((32, 2))
            for lane in range(32):
                tau = s - lag - (j0 + lane)
                c = self.cell(w, lane, tau)
                if c is not None:                                        # converters: diagonal read of the window
                    assert self.raw_row[w, tau & (RW - 1), lane] == tau, "window row overwritten or not loaded"
                    fac[lane] = self.raw[w, tau & (RW - 1), lane]
            if self.dir == 0:
                v = np.logaddexp(self.own[w], inn)
                self.own[w] = v + fac[:, 0]
                self.share[w] = v + fac[:, 1]
            else:
                v = np.logaddexp(self.own[w] + fac[:, 0], inn + fac[:, 1])
                self.own[w] = v
                self.share[w] = v.copy()
            self.val[w, blk % STAGES, k] = v
            self.edge[(s - lag) & (EDGE - 1), w] = self.share[w, 31]
            self.edge_tag[(s - lag) & (EDGE - 1), w] = s - lag

    # ---- consumer: value block vb of warp w -> output window -> rows; forwards the band boundary
    def consume(self, w, vb, plane, next_band):
        lag = self.lag(w)
        j0 = (self.band * self.nw + w) * 32
        for k in range(KB):
            for lane in range(32):
                tau = vb * KB + k - lag - (j0 + lane)
                self.out[w, tau & (WIN - 1), lane] = self.val[w, vb % STAGES, k, lane]
                self.out_row[w, tau & (WIN - 1), lane] = tau
        if next_band is not None and w == self.nw - 1:
            for k in range(KB):
                q = vb * KB + k - lag
                if q >= 0:
                    assert self.edge_tag[q & (EDGE - 1), w] == q, "boundary value overwritten before it was forwarded"
                    next_band.xedge[q] = self.edge[q & (EDGE - 1), w]
            next_band.xdone = vb + 1
        for k in range(KB):
            self.store_row(w, self.row_st[w] + k, plane)
        self.row_st[w] += KB
        self.consumed[w] = vb + 1

    def store_row(self, w, r, plane):
        for lane in range(32):
            c = self.cell(w, lane, r)
            if c is not None:
                assert self.out_row[w, r & (WIN - 1), lane] == r, "output row stored before it was complete"
                plane[c] = self.out[w, r & (WIN - 1), lane]


def sweep(lp2, Tb, Ub, nw, direction):
    """alpha (direction 0) or beta (1) plane of one utterance with bands of nw chain warps."""
    n_on = (Ub + 32) // 32
    n_bands = (n_on + nw - 1) // nw
    max_lag = (n_on - 1) * KB + ((n_on - 1) // nw) * (SKEW - KB)
    n_blocks = (Tb + Ub + max_lag + KB - 1) // KB
    plane = np.full((Tb, Ub + 1), np.nan)
    bands = [Band(b, n_bands, nw, direction, lp2, Tb, Ub) for b in range(n_bands)]
    for blk in range(n_blocks):
        for bi, band in enumerate(bands):
            nxt = bands[bi + 1] if bi + 1 < n_bands else None
            for w in range(nw):                       # warp w runs a whole block before warp w+1 starts it
                band.load_ahead(w, blk, n_blocks)     # loader as far ahead as its throttle allows
                if blk >= STAGES:                     # consumer as far BEHIND as the value ring allows
                    while band.consumed[w] < blk - STAGES + 1:
                        band.consume(w, band.consumed[w], plane, nxt)
                band.chain_block(w, blk, bands[bi - 1] if bi else None)
    for bi, band in enumerate(bands):
        nxt = bands[bi + 1] if bi + 1 < n_bands else None
        for w in range(nw):
            while band.consumed[w] < n_blocks:
                band.consume(w, band.consumed[w], plane, nxt)
            for r in range(band.row_st[w], Tb):       # rows the last lanes finished in the final blocks
                band.store_row(w, r, plane)
    return plane


@pytest.mark.parametrize("T,U,nw", [(40, 20, 1), (37, 70, 3), (90, 95, 3), (21, 130, 2), (70, 150, 2), (33, 200, 3),
                                    (5, 40, 2), (150, 3, 1)])
def test_schedule_model_reproduces_the_oracle(T, U, nw):
    rng = np.random.default_rng(T * 1000 + U)
    V = 5
    logits = rng.normal(size=(1, T, U + 1, V)) * 2.0
    labels = rng.integers(1, V, size=(1, U)).astype(np.int32)
    ref = np_oracle.rnnt_loss_np(logits, labels, [T], [U], 0)
    lp = log_softmax(logits[0])
    lab = np.concatenate([labels[0], [0]])
    lp2 = np.stack([lp[..., 0], np.take_along_axis(lp, lab[None, :, None].repeat(T, 0), 2)[..., 0]], -1)
    alpha = sweep(lp2, T, U, nw, 0)
    beta = sweep(lp2, T, U, nw, 1)
    assert not np.isnan(alpha).any() and not np.isnan(beta).any(), "a lattice cell was never stored"
    np.testing.assert_allclose(alpha, ref["alphas"][0], rtol=0, atol=1e-9)
    np.testing.assert_allclose(beta, ref["betas"][0], rtol=0, atol=1e-9)
    np.testing.assert_allclose(-beta[0, 0], ref["costs"][0], rtol=1e-12)
