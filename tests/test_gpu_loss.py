"""GPU parity: the dense-logits RNNTLoss drop-in and the lattice sweeps, through the C ABI, against
the CPU oracle (oracle/warp_cpu.c) and the committed golden vectors (torchaudio CPU).

Tolerances are the north_star's: fp32 per-utterance loss within 1e-5 relative, gradients within
1e-4 absolute.  fp16 / bf16 *inputs* (the reference's --precision=16 path, model.py:28-31) get a
separately stated looser bound because the logits themselves are rounded.
"""
import numpy as np
import pytest
import torch

import rnntransducer_b200 as rb
from conftest import load_golden
from rnntransducer_b200 import _lib, synthetic

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-5
GRAD_ATOL = 1e-4
DENSE = ["kat1.npz", "dense_full.npz", "dense_ragged.npz", "dense_blank_last.npz", "dense_v73.npz"]


def dev(x, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x
    t = t.cuda()
    return t.to(dtype) if dtype is not None else t


def oracle_pair(oracle_lib, d, blank=0):
    """The CPU oracle in fp32 (warp-transducer's precision) and in fp64 (the same algorithm without
    rounding).  On long lattices the fp32 oracle's own gradients drift from the fp64 ones by more
    than 1e-4 (alpha/beta reach |1e3|, ulp 6e-5, and are cancelled against each other), so the gate
    is: within GRAD_ATOL of the fp64 oracle, and within GRAD_ATOL + (fp32 oracle's own error) of
    the fp32 oracle."""
    a = [d["logits"].numpy() if torch.is_tensor(d["logits"]) else d["logits"]]
    a += [np.asarray(d[k]) for k in ("labels", "act_lens", "label_lens")]
    r32 = oracle_lib.rnnt_loss_cpu(*a, blank)
    r64 = oracle_lib.rnnt_loss_cpu(*a, blank, dtype=np.float64)
    return r32, r64


def check_against_oracles(costs, grads, r32, r64):
    np.testing.assert_allclose(costs, r64["costs"], rtol=LOSS_RTOL)
    np.testing.assert_allclose(costs, r32["costs"], rtol=LOSS_RTOL)
    np.testing.assert_allclose(grads, r64["grads"], atol=GRAD_ATOL)
    own = float(np.abs(r32["grads"] - r64["grads"]).max())
    np.testing.assert_allclose(grads, r32["grads"], atol=GRAD_ATOL + own)


def run_dense(logits, labels, act_lens, label_lens, blank=0, dtype=torch.float32):
    x = dev(logits, dtype).requires_grad_(True)
    costs = rb.rnnt_costs(x, dev(labels), dev(act_lens), dev(label_lens), blank)
    costs.sum().backward()
    return costs.detach().float().cpu().numpy(), x.grad.float().cpu().numpy()


@pytest.mark.parametrize("name", DENSE)
def test_dense_loss_matches_golden(cuda_lib, name):
    g = load_golden(name)
    costs, grads = run_dense(g["logits"], g["labels"], g["act_lens"], g["label_lens"], int(g["blank"]))
    np.testing.assert_allclose(costs, g["costs"], rtol=LOSS_RTOL)
    np.testing.assert_allclose(grads, g["grads"], atol=GRAD_ATOL)


def test_kat1_known_answer(cuda_lib):
    g = load_golden("kat1.npz")
    costs, _ = run_dense(g["logits"], g["labels"], g["act_lens"], g["label_lens"], 0)
    assert abs(float(costs[0]) - 4.49566698) < 4.5e-5


@pytest.mark.parametrize("ragged", [False, True])
@pytest.mark.parametrize("shape", [(4, 100, 20, 73), (3, 37, 40, 29), (2, 9, 70, 130), (1, 300, 1, 5)])
def test_dense_loss_matches_c_oracle(cuda_lib, oracle_lib, shape, ragged):
    B, T, U, V = shape
    d = synthetic.make_dense_logits(B, T, U, V, ragged=ragged, seed=100 + T)
    r32, r64 = oracle_pair(oracle_lib, d)
    costs, grads = run_dense(d["logits"], d["labels"], d["act_lens"], d["label_lens"])
    check_against_oracles(costs, grads, r32, r64)


def test_dense_edge_cases_match_oracle(cuda_lib, oracle_lib):
    """U_b = 0, T_b = 1, U = 0 for the whole batch (U1 = 1), label == blank id never drawn but
    blank != 0, and extreme logits (saturated softmax)."""
    d = synthetic.make_dense_logits(5, 13, 6, 11, ragged=True, seed=7)
    d["label_lens"][1] = 0
    d["act_lens"][2] = 1
    d["label_lens"][3], d["act_lens"][3] = 6, 1
    d["logits"][4] *= 30.0  # near one-hot softmax rows
    ref = oracle_lib.rnnt_loss_cpu(d["logits"].numpy(), d["labels"].numpy(), d["act_lens"].numpy(),
                                   d["label_lens"].numpy(), 0)
    costs, grads = run_dense(d["logits"], d["labels"], d["act_lens"], d["label_lens"])
    np.testing.assert_allclose(costs, ref["costs"], rtol=LOSS_RTOL)
    np.testing.assert_allclose(grads, ref["grads"], atol=GRAD_ATOL)
    # whole batch without labels: U1 == 1, labels tensor is [B, 0]
    x = torch.randn(3, 8, 1, 6, generator=torch.Generator().manual_seed(1))
    lab = torch.zeros(3, 0, dtype=torch.int32)
    al = torch.tensor([8, 3, 1], dtype=torch.int32)
    ll = torch.zeros(3, dtype=torch.int32)
    ref = oracle_lib.rnnt_loss_cpu(x.numpy(), lab.numpy(), al.numpy(), ll.numpy(), 2)
    costs, grads = run_dense(x, lab, al, ll, blank=2)
    np.testing.assert_allclose(costs, ref["costs"], rtol=LOSS_RTOL)
    np.testing.assert_allclose(grads, ref["grads"], atol=GRAD_ATOL)


def test_dense_properties(cuda_lib):
    """Oracle behaviours probed on the reference's loss: padded-region grads are
    exact zeros, sum_v grad = 0 per valid cell, mean == none.mean(), shapes per reduction."""
    d = synthetic.make_dense_logits(6, 50, 12, 73, ragged=True, seed=9)
    costs, grads = run_dense(d["logits"], d["labels"], d["act_lens"], d["label_lens"])
    for b in range(6):
        Tb, Ub = int(d["act_lens"][b]), int(d["label_lens"][b])
        assert np.all(grads[b, Tb:] == 0) and np.all(grads[b, :, Ub + 1:] == 0)
        assert np.abs(grads[b, :Tb, :Ub + 1].sum(-1)).max() < 2e-5
    args = [dev(d[k]) for k in ("logits", "labels", "act_lens", "label_lens")]
    mean = rb.RNNTLoss(0, "mean")(*args)
    ssum = rb.RNNTLoss(0, "sum")(*args)
    none = rb.RNNTLoss(0, "none")(*args)
    assert mean.shape == (1,) and ssum.shape == (1,) and none.shape == (6,)
    assert rb.RNNTLoss(0, "mean", warp_compat=False)(*args).dim() == 0
    np.testing.assert_allclose(float(mean), costs.mean(), rtol=1e-6)
    np.testing.assert_allclose(float(ssum), costs.sum(), rtol=1e-6)
    # functional form, warp-transducer argument order (north_star)
    f = rb.rnnt_loss(*args, 0, "mean")
    assert torch.equal(f, mean)


def test_dense_grad_scales_with_upstream_gradient(cuda_lib):
    d = synthetic.make_dense_logits(3, 20, 5, 17, ragged=True, seed=3)
    x = dev(d["logits"]).requires_grad_(True)
    args = [dev(d[k]) for k in ("labels", "act_lens", "label_lens")]
    w = torch.tensor([0.5, -2.0, 3.0], device="cuda")
    (rb.rnnt_costs(x, *args) * w).sum().backward()
    g1 = x.grad.clone()
    x.grad = None
    rb.rnnt_costs(x, *args).sum().backward()
    torch.testing.assert_close(g1, x.grad * w[:, None, None, None], atol=1e-6, rtol=1e-5)


@pytest.mark.parametrize("dtype,loss_rtol,grad_atol", [(torch.float16, 1e-3, 2e-3),
                                                       (torch.bfloat16, 1e-2, 1e-2)])
def test_dense_half_inputs(cuda_lib, oracle_lib, dtype, loss_rtol, grad_atol):
    """fp16 logits are what the reference's shipped script feeds torchaudio (run_train.sh:32).
    The oracle sees the SAME rounded logits, so the bound covers only our fp32 internals plus the
    rounding of the returned gradient to the input dtype (fp16: 2^-11, bf16: 2^-8 relative)."""
    d = synthetic.make_dense_logits(3, 40, 9, 73, ragged=True, seed=21)
    rounded = d["logits"].to(dtype).float()
    ref = oracle_lib.rnnt_loss_cpu(rounded.numpy(), d["labels"].numpy(), d["act_lens"].numpy(),
                                   d["label_lens"].numpy(), 0)
    costs, grads = run_dense(rounded, d["labels"], d["act_lens"], d["label_lens"], dtype=dtype)
    np.testing.assert_allclose(costs, ref["costs"], rtol=loss_rtol)
    np.testing.assert_allclose(grads, ref["grads"], atol=grad_atol)


def test_lattice_sweep_abi_alpha_beta(cuda_lib, oracle_lib):
    """rnntb200_lattice_sweep called directly: alpha / beta planes against the oracle's, costs
    against -beta(0,0), and the alpha-side log-likelihood cross-check."""
    B, T, U, V = 5, 61, 70, 9  # U1 = 71 > 64: exercises the multi-warp hand-off
    d = synthetic.make_dense_logits(B, T, U, V, ragged=True, seed=5)
    ref = oracle_lib.rnnt_loss_cpu(d["logits"].numpy(), d["labels"].numpy(), d["act_lens"].numpy(),
                                   d["label_lens"].numpy(), 0, want_alpha_beta=True)
    lp = torch.log_softmax(d["logits"], -1)
    U1 = U + 1
    lab = torch.cat([d["labels"].long(), torch.zeros(B, 1, dtype=torch.long)], 1)
    lp_label = lp.gather(3, lab[:, None, :, None].expand(B, T, U1, 1))[..., 0]
    lp2 = torch.stack([lp[..., 0], lp_label], -1).contiguous().cuda()
    al, ll = dev(d["act_lens"]), dev(d["label_lens"])
    f32 = dict(device="cuda", dtype=torch.float32)
    alpha, beta = (torch.zeros(B, T, U1, device="cuda", dtype=torch.int32) for _ in range(2))  # e16m16
    costs, ll_alpha = torch.zeros(B, **f32), torch.zeros(B, **f32)
    st = cuda_lib.rnntb200_lattice_sweep(lp2.data_ptr(), al.data_ptr(), ll.data_ptr(), B, T, U1,
                                         alpha.data_ptr(), beta.data_ptr(), costs.data_ptr(),
                                         ll_alpha.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _lib.check(st, "rnntb200_lattice_sweep")
    torch.cuda.synchronize()
    np.testing.assert_allclose(costs.cpu().numpy(), ref["costs"], rtol=LOSS_RTOL)
    np.testing.assert_allclose(-ll_alpha.cpu().numpy(), ref["costs"], rtol=LOSS_RTOL)
    # planes are e16m16 wide-exponent floats (include/rnnt_b200.h)
    def to_ln(q):
        q = q.cpu().numpy().astype(np.int64)
        return ((q >> 16) + np.log2(1.0 + (q & 0xFFFF) / 65536.0)) * np.log(2.0)
    ref64 = oracle_lib.rnnt_loss_cpu(d["logits"].numpy(), d["labels"].numpy(), d["act_lens"].numpy(),
                                     d["label_lens"].numpy(), 0, want_alpha_beta=True, dtype=np.float64)
    for b in range(B):
        Tb, Ub = int(d["act_lens"][b]), int(d["label_lens"][b])
        np.testing.assert_allclose(to_ln(alpha[b, :Tb, :Ub + 1]), ref64["alphas"][b, :Tb, :Ub + 1], atol=5e-5)
        np.testing.assert_allclose(to_ln(beta[b, :Tb, :Ub + 1]), ref64["betas"][b, :Tb, :Ub + 1], atol=5e-5)
        # the fp32 oracle agrees to its own rounding (ulp(|alpha|) per step)
        np.testing.assert_allclose(to_ln(alpha[b, :Tb, :Ub + 1]), ref["alphas"][b, :Tb, :Ub + 1],
                                   rtol=2e-5, atol=2e-4)
        # cells outside the utterance's box are left untouched by forward calls (header contract)
        assert int(alpha[b, Tb:].abs().sum()) == 0 and int(beta[b, :, Ub + 1:].abs().sum()) == 0


@pytest.mark.parametrize("U", [0, 1, 30, 31, 32, 62, 63, 64, 94, 95, 96, 126, 127, 128, 160, 200, 520, 800])
def test_sweep_warp_boundaries(cuda_lib, oracle_lib, U):
    """Label lengths around every multiple of 32 and in every regime of the sweep dispatch: one to four warps
    of the self-contained sweep in one CTA, bands of two / three / four warps in a thread-block cluster beyond
    (U1 <= 512 / 768 / 1024); T shorter and longer than the FIFOs, ragged lengths included.  (The round's GPU
    visits ran the same test under RNNTB200_SWEEP=ws and RNNTB200_SWEEP_BW=2,3,4 as well.)"""
    for T in (3, 45, 150):
        d = synthetic.make_dense_logits(3, T, U, 6, ragged=True, seed=100 + U + T)
        r32, r64 = oracle_pair(oracle_lib, d)
        costs, grads = run_dense(d["logits"], d["labels"], d["act_lens"], d["label_lens"])
        check_against_oracles(costs, grads, r32, r64)


def test_sweep_many_utterances(cuda_lib, oracle_lib):
    """Several waves of CTAs (B = 400 utterances, two warps each): every utterance against the oracle, and the
    fused mean over the batch against the sum of the per-utterance costs."""
    d = synthetic.make_dense_logits(400, 24, 40, 5, ragged=True, seed=77)
    r32, r64 = oracle_pair(oracle_lib, d)
    costs, grads = run_dense(d["logits"], d["labels"], d["act_lens"], d["label_lens"])
    check_against_oracles(costs, grads, r32, r64)


def test_out_of_range_lengths_and_labels_are_clamped_identically(cuda_lib):
    """Nobody validates device-side lengths / labels on the hot path; every kernel clamps them the same way
    (csrc/common.cuh), so garbage gives the result of the clamped inputs -- finite, no out-of-bounds access --
    and RNNTLoss(check_lengths=True) raises instead."""
    d = synthetic.make_batch(4, 30, 7, 11, 128, ragged=True, seed=8, device="cuda")
    bad = {k: v.clone() for k, v in d.items()}
    bad["act_lens"][1], bad["act_lens"][2] = 0, 1000       # -> 1, T
    bad["label_lens"][0], bad["label_lens"][3] = -5, 99     # -> 0, U
    bad["labels"][2, 0], bad["labels"][2, 1] = 500, -3      # -> V-1, 0
    good = {k: v.clone() for k, v in d.items()}
    good["act_lens"][1], good["act_lens"][2] = 1, 30
    good["label_lens"][0], good["label_lens"][3] = 0, 7
    good["labels"][2, 0], good["labels"][2, 1] = 10, 0
    run = lambda x: rb.joint_rnnt_costs(x["enc"], x["dec"], x["weight"], x["bias"], x["labels"], x["act_lens"],
                                        x["label_lens"], 0, deterministic=True)
    a, b = run(bad), run(good)
    assert torch.isfinite(a).all() and torch.equal(a, b)
    logits = rb.joint_dense(d["enc"], d["dec"], d["weight"], d["bias"])
    assert torch.equal(rb.rnnt_costs(logits, bad["labels"], bad["act_lens"], bad["label_lens"]),
                       rb.rnnt_costs(logits, good["labels"], good["act_lens"], good["label_lens"]))
    with pytest.raises(RuntimeError, match="out of range|labels must be"):
        rb.RNNTLoss(0, "mean", check_lengths=True)(logits, bad["labels"], bad["act_lens"], bad["label_lens"])


def test_long_lattice_matches_oracle(cuda_lib, oracle_lib):
    """cfg-3-shaped deep sweep (T=1500, U=300; 1800 anti-diagonals) at small V."""
    d = synthetic.make_dense_logits(2, 1500, 300, 8, ragged=True, seed=33)
    r32, r64 = oracle_pair(oracle_lib, d)
    costs, grads = run_dense(d["logits"], d["labels"], d["act_lens"], d["label_lens"])
    check_against_oracles(costs, grads, r32, r64)


def test_full_size_cfg2_dense_properties(cuda_lib):
    """BASELINE cfg 2 at full size (B=32,T=400,U=80,V=73; 303 MB of logits) through
    size-independent properties: finite costs, alpha- and beta-side likelihoods agree,
    sum_v grad = 0, padded grads exactly zero, and a sub-batch agrees with the full batch."""
    c = synthetic.CONFIGS[2]
    d = synthetic.make_dense_logits(c["B"], c["T"], c["U"], c["V"], ragged=True, seed=1236, device="cuda")
    x = d["logits"].requires_grad_(True)
    costs = rb.rnnt_costs(x, d["labels"], d["act_lens"], d["label_lens"])
    costs.sum().backward()
    assert torch.isfinite(costs).all() and float(costs.min()) > 0
    g = x.grad
    assert float(g.sum(-1).abs().max()) < 5e-5
    t_idx = torch.arange(c["T"], device="cuda")[None, :, None]
    u_idx = torch.arange(c["U"] + 1, device="cuda")[None, None, :]
    pad = (t_idx >= d["act_lens"][:, None, None]) | (u_idx > d["label_lens"][:, None, None])
    assert float(g[pad].abs().max()) == 0.0
    sub = slice(5, 9)
    c_sub = rb.rnnt_costs(x.detach()[sub].contiguous(), d["labels"][sub].contiguous(),
                          d["act_lens"][sub].contiguous(), d["label_lens"][sub].contiguous())
    torch.testing.assert_close(c_sub, costs.detach()[sub], rtol=1e-6, atol=0)


def test_cuda_graph_capture_of_dense_fwd_bwd(cuda_lib):
    """Every entry point only enqueues work on the caller's stream (header contract), so a whole
    fwd+bwd is CUDA-graph capturable and replays bit-identically."""
    d = synthetic.make_dense_logits(4, 30, 7, 19, ragged=True, seed=2, device="cuda")
    x = d["logits"].clone().requires_grad_(True)
    args = (d["labels"], d["act_lens"], d["label_lens"])
    rb.rnnt_costs(x, *args).sum().backward()  # warm-up outside capture
    eager = x.grad.clone()
    x.grad = None
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        rb.rnnt_costs(x, *args).sum().backward()
        x.grad = None
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        rb.rnnt_costs(x, *args).sum().backward()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(x.grad, eager)


def test_no_cpu_fallback(cuda_lib):
    d = synthetic.make_dense_logits(2, 5, 3, 7, seed=1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rb.rnnt_costs(d["logits"], d["labels"], d["act_lens"], d["label_lens"])
    with pytest.raises(RuntimeError, match="same device"):
        rb.rnnt_costs(d["logits"].cuda(), d["labels"], d["act_lens"].cuda(), d["label_lens"].cuda())
