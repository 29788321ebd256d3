"""GPU parity: fused joint + RNN-T loss in the reference's own joint (concat -> GELU(tanh) ->
Linear(2H -> V), networks/transducer.py:54-71), evaluated in factorised form without ever forming
the [B,T,U1,V] logits.  Checked against
  - golden vectors produced by the reference's JointNet.joint + torchaudio CPU loss + autograd
    (oracle/gen_golden.py),
  - the CPU restatement oracle/joint_ref.py + oracle/warp_cpu.c on seeded inputs,
  - our own dense path at BASELINE cfg-2 size (size-independent cross-check).
Tolerances: loss 1e-5 relative, gradients 1e-4 absolute (north_star).  Parameter gradients
(d_weight, d_bias) are sums of the per-cell gradients over every lattice cell of the batch and
reach |1e2|; for those the bound is conftest.param_atol (1e-4 + 5e-5 * max|ref|), see there.
"""
import numpy as np
import pytest
import torch

import rnntransducer_b200 as rb
from conftest import load_golden, param_atol
from oracle import joint_ref
from rnntransducer_b200 import synthetic

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-5
GRAD_ATOL = 1e-4


def to_cuda(d):
    return {k: (torch.from_numpy(np.ascontiguousarray(v)) if isinstance(v, np.ndarray) else v).cuda()
            for k, v in d.items() if isinstance(v, (np.ndarray, torch.Tensor)) and np.ndim(v) > 0}


def ref_step(d, mode="concat_gelu", dtype=torch.float32):
    return joint_ref.joint_loss_fwd_bwd(d["enc"], d["dec"], d["weight"], d["bias"], d["labels"].numpy(),
                                        d["act_lens"].numpy(), d["label_lens"].numpy(), 0, "mean", mode,
                                        dtype=dtype)


def fused_step(d, reduction="mean", deterministic=False, mode="concat_gelu", gemm="fp32"):
    t = {k: d[k].clone().requires_grad_(True) for k in ("enc", "dec", "weight", "bias")}
    costs = rb.joint_rnnt_costs(t["enc"], t["dec"], t["weight"], t["bias"], d["labels"], d["act_lens"],
                                d["label_lens"], 0, mode, gemm, deterministic)
    loss = {"mean": costs.mean(), "sum": costs.sum()}[reduction]
    loss.backward()
    out = dict(costs=costs.detach().cpu().numpy(), loss=float(loss))
    out.update({"d_" + k: v.grad.cpu().numpy() for k, v in t.items()})
    return out


@pytest.mark.parametrize("deterministic", [False, True])
@pytest.mark.parametrize("name", ["joint_small_full.npz", "joint_small_ragged.npz"])
def test_fused_matches_reference_jointnet_golden(cuda_lib, name, deterministic):
    g = load_golden(name)
    r = fused_step(to_cuda(g), deterministic=deterministic)
    np.testing.assert_allclose(r["costs"], g["costs"], rtol=LOSS_RTOL)
    np.testing.assert_allclose(r["loss"], float(g["loss"]), rtol=LOSS_RTOL)
    for k in ("d_enc", "d_dec", "d_weight", "d_bias"):
        np.testing.assert_allclose(r[k], g[k], atol=GRAD_ATOL, err_msg=k)


@pytest.mark.parametrize("name,ragged", [("joint_cfg1.npz", False), ("joint_cfg1_ragged.npz", True)])
def test_fused_cfg1_matches_reference_golden(cuda_lib, name, ragged):
    """BASELINE cfg 1 (B=4,T=100,U=20,V=73,H=320): inputs regenerated from the seed, outputs are
    the sub-sampled reference results stored by gen_golden.py."""
    g = load_golden(name)
    c = synthetic.CONFIGS[1]
    d = synthetic.make_batch(c["B"], c["T"], c["U"], c["V"], c["H"], ragged=ragged, seed=int(g["seed"]),
                             device="cuda")
    r = fused_step(d)
    st, sh = int(g["stride_t"]), int(g["stride_h"])
    np.testing.assert_allclose(r["costs"], g["costs"], rtol=LOSS_RTOL)
    np.testing.assert_allclose(r["d_bias"], g["d_bias"], atol=param_atol(g["d_bias"]))
    np.testing.assert_allclose(r["d_weight"][:, ::sh], g["d_weight"], atol=param_atol(g["d_weight"]))
    np.testing.assert_allclose(r["d_enc"][:, ::st, ::sh], g["d_enc"], atol=GRAD_ATOL)
    np.testing.assert_allclose(r["d_dec"][:, :, ::sh], g["d_dec"], atol=GRAD_ATOL)


@pytest.mark.parametrize("shape", [(3, 33, 9, 73, 48), (2, 20, 40, 130, 24), (5, 8, 3, 5, 16)])
def test_fused_matches_cpu_restatement(cuda_lib, oracle_lib, shape):
    B, T, U, V, H = shape
    d = synthetic.make_batch(B, T, U, V, H, ragged=True, seed=77 + V)
    if B >= 3:
        d["label_lens"][1] = 0
        d["labels"][1] = 0
        d["act_lens"][2] = 1
    # fp64 run of the same restatement: the fp32 reference's own rounding on d_weight (a sum over
    # every lattice cell) is not negligible against 1e-4, so the gate is 1e-4 against fp64 and
    # 1e-4 + (fp32 reference's own error) against fp32
    ref, ref64 = ref_step(d), ref_step(d, dtype=torch.float64)
    for det in (False, True):
        r = fused_step({k: v.cuda() for k, v in d.items()}, deterministic=det)
        np.testing.assert_allclose(r["costs"], ref["costs"], rtol=LOSS_RTOL)
        for k in ("d_enc", "d_dec", "d_weight", "d_bias"):
            np.testing.assert_allclose(r[k], ref64[k], atol=GRAD_ATOL, err_msg=k)
            own = float(np.abs(ref[k] - ref64[k]).max())
            np.testing.assert_allclose(r[k], ref[k], atol=GRAD_ATOL + own, err_msg=k)


@pytest.mark.parametrize("scale", [8.0, 40.0, 150.0])
def test_fused_extreme_logit_ranges(cuda_lib, oracle_lib, scale):
    """exp(P_enc + P_dec) is evaluated as exp(P_enc) * exp(P_dec) (joint_cg_mm.cu).  With logit ranges
    of tens to hundreds of nats in BOTH projections and peaks that do not line up, the factorised
    partition underflows and cells must take the exact path; results still match the oracle."""
    d = synthetic.make_batch(3, 21, 7, 11, 16, ragged=True, seed=321)
    d["weight"] = d["weight"] * scale
    ref64 = ref_step(d, dtype=torch.float64)
    for det in (False, True):
        r = fused_step({k: v.cuda() for k, v in d.items()}, deterministic=det)
        np.testing.assert_allclose(r["costs"], ref64["costs"], rtol=2e-5)
        for k in ("d_enc", "d_dec", "d_weight", "d_bias"):
            np.testing.assert_allclose(r[k], ref64[k], atol=param_atol(ref64[k], 1e-4 * scale), err_msg=k)


@pytest.mark.parametrize("shape", [(3, 50, 9, 73, 64, 64), (2, 130, 20, 73, 512, 512), (1, 7, 3, 40, 128, 320)])
def test_tensor_core_projection_matches_fp64(cuda_lib, shape):
    """rnntb200_joint_cg_project (bf16 hi/lo split, 3 tcgen05 MMAs per product) against an fp64
    evaluation of gelu_tanh(x) W^T + b: fp32-class accuracy (2e-5 absolute on O(1) outputs), and its
    backward against autograd through the fp64 expression."""
    from rnntransducer_b200.loss import project_concat_gelu
    B, T, U, V, He, Hd = shape
    g = torch.Generator().manual_seed(5)
    enc = torch.randn(B, T, He, generator=g).cuda().requires_grad_(True)
    dec = torch.randn(B, U + 1, Hd, generator=g).cuda().requires_grad_(True)
    w = ((torch.rand(V, He + Hd, generator=g) - 0.5) * 0.1).cuda().requires_grad_(True)
    b = ((torch.rand(V, generator=g) - 0.5) * 0.1).cuda().requires_grad_(True)
    assert cuda_lib.rnntb200_joint_cg_project_workspace_bytes(V, He, Hd) > 0
    penc, pdec = project_concat_gelu(enc, dec, w, b)
    up_e = torch.randn(penc.shape, generator=g).cuda()
    up_d = torch.randn(pdec.shape, generator=g).cuda()
    ((penc * up_e).sum() + (pdec * up_d).sum()).backward()
    e64, d64, w64, b64 = (t.detach().double().requires_grad_(True) for t in (enc, dec, w, b))
    gelu = lambda x: torch.nn.functional.gelu(x, approximate="tanh")
    re = gelu(e64) @ w64[:, :He].T + b64
    rd = gelu(d64) @ w64[:, He:].T
    ((re * up_e.double()).sum() + (rd * up_d.double()).sum()).backward()
    torch.testing.assert_close(penc.double(), re, atol=2e-5, rtol=0)
    torch.testing.assert_close(pdec.double(), rd, atol=2e-5, rtol=0)
    for ours, ref in ((enc, e64), (dec, d64), (w, w64), (b, b64)):
        torch.testing.assert_close(ours.grad.double(), ref.grad, atol=param_atol(ref.grad.cpu().numpy()), rtol=0)


def test_deterministic_mode_is_bit_reproducible(cuda_lib):
    d = synthetic.make_batch(4, 64, 17, 73, 32, ragged=True, seed=5, device="cuda")
    a = fused_step(d, deterministic=True)
    b = fused_step(d, deterministic=True)
    for k in ("costs", "d_enc", "d_dec", "d_weight", "d_bias"):
        assert np.array_equal(a[k], b[k]), k


def test_full_size_cfg2_fused_vs_dense_path(cuda_lib):
    """BASELINE cfg 2 at full size: the fused path (no logits in HBM) against our dense-logits path
    fed by the eager joint (303 MB logits), plus size-independent properties."""
    c = synthetic.CONFIGS[2]
    d = synthetic.make_batch(c["B"], c["T"], c["U"], c["V"], c["H"], ragged=True, seed=1236, device="cuda")
    fused = fused_step(d)
    t = {k: d[k].clone().requires_grad_(True) for k in ("enc", "dec", "weight", "bias")}
    logits = rb.joint_dense(t["enc"], t["dec"], t["weight"], t["bias"])
    costs = rb.rnnt_costs(logits, d["labels"], d["act_lens"], d["label_lens"])
    costs.mean().backward()
    np.testing.assert_allclose(fused["costs"], costs.detach().cpu().numpy(), rtol=LOSS_RTOL)
    for k in ("enc", "dec", "weight", "bias"):
        ref = t[k].grad.cpu().numpy()
        np.testing.assert_allclose(fused["d_" + k], ref, err_msg=k,
                                   atol=param_atol(ref) if k in ("weight", "bias") else GRAD_ATOL)
    # sum_v g = 0 per cell  =>  the bias gradient sums to zero
    assert abs(float(fused["d_bias"].sum())) < param_atol(fused["d_bias"])
    # frames past an utterance's length receive exactly zero gradient
    al = d["act_lens"].cpu().numpy()
    for b in range(c["B"]):
        assert np.all(fused["d_enc"][b, al[b]:] == 0)


def test_jointnet_module_end_to_end(cuda_lib):
    """The reference call shape (model.py:49,56-57): logits = net(...); loss = RNNTLoss(...)(logits,
    targets, tensor_audio_lengths, target_lengths); loss.backward().  Fused (lazy handle) against
    fused=False (dense logits into the same loss)."""
    g = load_golden("jointnet_fwd.npz")
    ep = dict(input_size=8, hidden_size=12, output_size=16, num_layers=2, rnn_type="gru",
              dropout=0.0, bidirectional=True)
    dp = dict(embedding_size=11, pad_token_id=0, hidden_size=12, output_size=16, num_layers=2,
              rnn_type="lstm", dropout=0.0)
    sd = {k[len("sd__"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd__")}
    audio = torch.from_numpy(g["audio"]).cuda()
    texts = torch.from_numpy(g["texts"]).cuda()
    audio_lengths, text_lengths = g["audio_lengths"].tolist(), g["text_lengths"].tolist()
    targets = texts[:, 1:].to(torch.int32).contiguous()
    act_lens = torch.tensor(audio_lengths, dtype=torch.int32, device="cuda")
    label_lens = torch.tensor([n - 1 for n in text_lengths], dtype=torch.int32, device="cuda")
    loss_fn = rb.RNNTLoss(blank=0, reduction="mean")
    results = {}
    for fused in (True, False):
        net = rb.JointNet(dict(ep), dict(dp), 11, fused=fused).cuda().train()
        net.load_state_dict(sd)
        logits = net(audio, audio_lengths, texts, text_lengths)
        assert isinstance(logits, rb.JointLogits) == fused
        assert tuple(logits.shape) == (3, 14, 5, 11)
        loss = loss_fn(logits, targets, act_lens, label_lens)
        assert loss.shape == (1,)
        loss.backward()
        results[fused] = (float(loss), {n: p.grad.clone() for n, p in net.named_parameters()})
        if fused:  # the handle still materialises to the reference logits when asked
            np.testing.assert_allclose(logits.materialize().detach().cpu().numpy(), g["logits"], atol=1e-4)
    assert abs(results[True][0] - results[False][0]) < 1e-5 * abs(results[False][0])
    for n, gr in results[False][1].items():
        torch.testing.assert_close(results[True][1][n], gr, atol=1e-4, rtol=1e-3, msg=n)


def test_full_size_cfg3_long_utterances(cuda_lib):
    """BASELINE cfg 3 (B=8, T=1500, U=300, V=73, H=512; 1800 anti-diagonals, 10-warp sweeps): fused path
    against our dense-logits path on the same inputs (1.05 GB of logits), plus the size-independent
    properties."""
    c = synthetic.CONFIGS[3]
    d = synthetic.make_batch(c["B"], c["T"], c["U"], c["V"], c["H"], ragged=True, seed=1237, device="cuda")
    fused = fused_step(d)
    assert np.isfinite(fused["costs"]).all() and (fused["costs"] > 0).all()
    with torch.no_grad():
        logits = rb.joint_dense(d["enc"], d["dec"], d["weight"], d["bias"])
    logits.requires_grad_(True)
    costs = rb.rnnt_costs(logits, d["labels"], d["act_lens"], d["label_lens"])
    np.testing.assert_allclose(fused["costs"], costs.detach().cpu().numpy(), rtol=LOSS_RTOL)
    costs.mean().backward()
    g = logits.grad  # d loss / d logits: its reductions are the gradients of the two projections' bias
    d_bias_dense = g.sum((0, 1, 2)).cpu().numpy()
    np.testing.assert_allclose(fused["d_bias"], d_bias_dense, atol=param_atol(d_bias_dense))
    assert abs(float(fused["d_bias"].sum())) < param_atol(fused["d_bias"])
    al = d["act_lens"].cpu().numpy()
    for b in range(c["B"]):
        assert np.all(fused["d_enc"][b, al[b]:] == 0)


@pytest.mark.parametrize("generic", [False, True])
def test_large_vocab_cfg4_kernels(cuda_lib, monkeypatch, generic):
    """V = 1024 (BASELINE cfg 4's vocabulary): the wide factorised kernels (128-column chunks,
    joint_cg_mm.cu) and, with RNNTB200_CG_GENERIC set, the generic per-cell kernels (joint_cg.cu):
    fused against the dense path at B = 3."""
    if generic:
        monkeypatch.setenv("RNNTB200_CG_GENERIC", "1")
    c = synthetic.CONFIGS[4]
    d = synthetic.make_batch(3, c["T"], c["U"], c["V"], c["H"], ragged=True, seed=1238, device="cuda")
    assert (cuda_lib.rnntb200_joint_cg_factors_bytes(3, c["T"], c["U"] + 1, c["V"]) == 0) == generic
    fused = fused_step(d)
    t = {k: d[k].clone().requires_grad_(True) for k in ("enc", "dec", "weight", "bias")}
    logits = rb.joint_dense(t["enc"], t["dec"], t["weight"], t["bias"])
    costs = rb.rnnt_costs(logits, d["labels"], d["act_lens"], d["label_lens"])
    costs.mean().backward()
    np.testing.assert_allclose(fused["costs"], costs.detach().cpu().numpy(), rtol=LOSS_RTOL)
    for k in ("enc", "dec", "weight", "bias"):
        ref = t[k].grad.cpu().numpy()
        np.testing.assert_allclose(fused["d_" + k], ref, err_msg=k,
                                   atol=param_atol(ref) if k in ("weight", "bias") else GRAD_ATOL)


@pytest.mark.parametrize("reduction", ["mean", "sum"])
def test_fused_reduction_node_matches_costs_then_reduce(cuda_lib, reduction):
    """RNNTLoss(reduction=mean|sum) on the fused path is ONE autograd node (the sweep's last-arriving utterance
    adds up the costs, the gradient kernel scales the one upstream value): same loss and gradients as
    per-utterance costs followed by torch's reduction; the arrival counter re-arms itself (repeat calls and
    CUDA-graph replays give the same bits)."""
    d = synthetic.make_batch(5, 70, 13, 73, 128, ragged=True, seed=91, device="cuda")

    def run(fused):
        t = {k: d[k].clone().requires_grad_(True) for k in ("enc", "dec", "weight", "bias")}
        if fused:
            loss = rb.joint_rnnt_loss(t["enc"], t["dec"], t["weight"], t["bias"], d["labels"], d["act_lens"],
                                      d["label_lens"], 0, reduction, deterministic=True)
        else:
            costs = rb.joint_rnnt_costs(t["enc"], t["dec"], t["weight"], t["bias"], d["labels"], d["act_lens"],
                                        d["label_lens"], 0, deterministic=True)
            loss = (costs.mean() if reduction == "mean" else costs.sum()).reshape(1)
        (loss * 3.0).backward()  # a non-trivial upstream gradient
        return loss.detach(), {k: v.grad for k, v in t.items()}

    l0, g0 = run(False)
    l1, g1 = run(True)
    l2, g2 = run(True)
    assert l1.shape == (1,) and torch.equal(l1, l2)
    torch.testing.assert_close(l1, l0, rtol=1e-6, atol=0)
    for k in g0:
        assert torch.equal(g1[k], g2[k]) or k in ("weight", "bias"), k  # (projection backward sums tiles with atomics)
        torch.testing.assert_close(g1[k], g0[k], rtol=1e-5, atol=1e-6 * float(g0[k].abs().max()) + 1e-9, msg=k)
    # the module form takes the same route
    h = rb.JointLogits(d["enc"], d["dec"], d["weight"], d["bias"], "concat_gelu")
    lm = rb.RNNTLoss(0, reduction)(h, d["labels"], d["act_lens"], d["label_lens"])
    torch.testing.assert_close(lm, l1, rtol=1e-6, atol=0)
    # captured in a CUDA graph: replays re-arm the counter too
    enc, dec, w, b = (d[k].clone().requires_grad_(True) for k in ("enc", "dec", "weight", "bias"))
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        rb.joint_rnnt_loss(enc, dec, w, b, d["labels"], d["act_lens"], d["label_lens"], 0, reduction,
                           deterministic=True).backward()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    for t in (enc, dec, w, b):
        t.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        lg = rb.joint_rnnt_loss(enc, dec, w, b, d["labels"], d["act_lens"], d["label_lens"], 0, reduction,
                                deterministic=True)
        (lg * 3.0).backward()  # the same upstream gradient as above: the same bits
    for _ in range(3):
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(lg, l1)
        assert torch.equal(enc.grad, g1["enc"]) and torch.equal(dec.grad, g1["dec"])
