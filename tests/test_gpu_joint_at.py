"""GPU parity: fused joint + RNN-T loss in the north_star's joint, logits = fc(tanh(enc_t + dec_u)).

The reference has no such joint; the oracle of this mode is
torchaudio.models.rnnt._Joiner(activation="tanh") + torchaudio CPU loss (golden vectors) and the
restatement oracle/joint_ref.py + oracle/warp_cpu.c.

Tolerances
  gemm="fp32" (CUDA-core FFMA):  loss 1e-5 relative, gradients 1e-4 absolute (north_star).
  gemm="bf16" (tensor-core numerics: operands rounded to bf16, fp32 accumulation): the separately
      stated looser bound -- loss 5e-3 relative, gradients 2e-2 * max|grad| absolute.  Operand
      rounding is 2^-9 relative; a logit is a K = H term dot product, so its error is
      ~2^-9 * sqrt(H) * |w||z| ~ 1e-2 at H = 512, which moves per-cell log-probs by about that
      much and the loss (a sum over T+U cells of a path) by < 1e-3 relative.
"""
import numpy as np
import pytest
import torch

import rnntransducer_b200 as rb
from conftest import load_golden, param_atol
from oracle import joint_ref
from rnntransducer_b200 import synthetic
from test_gpu_joint_cg import fused_step, to_cuda

pytestmark = pytest.mark.gpu

KEYS = ("d_enc", "d_dec", "d_weight", "d_bias")


def check(r, ref, gemm):
    if gemm == "fp32":
        np.testing.assert_allclose(r["costs"], ref["costs"], rtol=1e-5)
        for k in KEYS:
            atol = param_atol(ref[k]) if k in ("d_weight", "d_bias") else 1e-4
            np.testing.assert_allclose(r[k], ref[k], atol=atol, err_msg=k)
    else:
        np.testing.assert_allclose(r["costs"], ref["costs"], rtol=5e-3)
        for k in KEYS:
            np.testing.assert_allclose(r[k], ref[k], atol=2e-2 * max(1e-3, float(np.abs(ref[k]).max())),
                                       err_msg=k)


@pytest.mark.parametrize("gemm", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["addtanh_small_full.npz", "addtanh_small_ragged.npz"])
def test_add_tanh_matches_torchaudio_joiner_golden(cuda_lib, name, gemm):
    g = load_golden(name)
    r = fused_step(to_cuda(g), mode="add_tanh", gemm=gemm)
    check(r, g, gemm)
    np.testing.assert_allclose(r["loss"], float(g["loss"]), rtol=1e-5 if gemm == "fp32" else 5e-3)


@pytest.mark.parametrize("gemm", ["fp32", "bf16"])
@pytest.mark.parametrize("shape", [(3, 33, 9, 73, 64), (2, 21, 18, 200, 96), (4, 9, 3, 5, 40),
                                   (2, 50, 12, 73, 512), (2, 20, 10, 300, 128), (3, 70, 20, 73, 256)])
def test_add_tanh_matches_cpu_restatement(cuda_lib, oracle_lib, shape, gemm):
    B, T, U, V, H = shape
    d = synthetic.make_batch(B, T, U, V, H, mode="add_tanh", ragged=True, seed=55 + V)
    if B >= 3:
        d["label_lens"][1] = 0
        d["labels"][1] = 0
        d["act_lens"][2] = 1
    ref = joint_ref.joint_loss_fwd_bwd(d["enc"], d["dec"], d["weight"], d["bias"], d["labels"].numpy(),
                                       d["act_lens"].numpy(), d["label_lens"].numpy(), 0, "mean", "add_tanh")
    r = fused_step({k: v.cuda() for k, v in d.items()}, mode="add_tanh", gemm=gemm)
    check(r, ref, gemm)


def test_add_tanh_properties_cfg2_slice(cuda_lib):
    """A cfg-2-shaped slice (T=400, U=80, V=73, H=512, B=2): fused path against our own dense path fed
    by the eager add-tanh joint, d_bias sums to zero, padded frames get exactly zero gradient."""
    d = synthetic.make_batch(2, 400, 80, 73, 512, mode="add_tanh", ragged=True, seed=1236, device="cuda")
    d["act_lens"][1] = 333
    fused = fused_step(d, mode="add_tanh", gemm="fp32")
    t = {k: d[k].clone().requires_grad_(True) for k in ("enc", "dec", "weight", "bias")}
    logits = rb.joint_dense(t["enc"], t["dec"], t["weight"], t["bias"], "add_tanh")
    costs = rb.rnnt_costs(logits, d["labels"], d["act_lens"], d["label_lens"])
    costs.mean().backward()
    np.testing.assert_allclose(fused["costs"], costs.detach().cpu().numpy(), rtol=1e-5)
    for k in ("enc", "dec", "weight", "bias"):
        ref = t[k].grad.cpu().numpy()
        np.testing.assert_allclose(fused["d_" + k], ref, err_msg=k,
                                   atol=param_atol(ref) if k in ("weight", "bias") else 1e-4)
    assert abs(float(fused["d_bias"].sum())) < param_atol(fused["d_bias"])
    assert np.all(fused["d_enc"][1, 333:] == 0)


def test_add_tanh_module_end_to_end(cuda_lib):
    ep = dict(input_size=8, hidden_size=12, output_size=32, num_layers=1, rnn_type="lstm",
              dropout=0.0, bidirectional=False)
    dp = dict(embedding_size=11, pad_token_id=0, hidden_size=12, output_size=32, num_layers=1,
              rnn_type="lstm", dropout=0.0)
    torch.manual_seed(3)
    net = rb.JointNet(ep, dp, 11, mode="add_tanh", gemm="fp32").cuda()
    audio = torch.randn(3, 14, 8, device="cuda")
    texts = torch.tensor([[0, 3, 4, 5, 1], [0, 2, 2, 0, 0], [0, 7, 8, 9, 0]], device="cuda")
    audio_lengths, text_lengths = [14, 9, 11], [5, 3, 4]
    targets = texts[:, 1:].to(torch.int32).contiguous()
    act_lens = torch.tensor(audio_lengths, dtype=torch.int32, device="cuda")
    label_lens = torch.tensor([n - 1 for n in text_lengths], dtype=torch.int32, device="cuda")
    loss_fn = rb.RNNTLoss(0, "mean")
    logits = net(audio, audio_lengths, texts, text_lengths)
    assert isinstance(logits, rb.JointLogits)
    loss = loss_fn(logits, targets, act_lens, label_lens)
    loss.backward()
    fused = {n: p.grad.clone() for n, p in net.named_parameters()}
    net.zero_grad()
    net.fused = False
    loss2 = loss_fn(net(audio, audio_lengths, texts, text_lengths), targets, act_lens, label_lens)
    loss2.backward()
    assert abs(float(loss) - float(loss2)) < 1e-5 * abs(float(loss2))
    for n, p in net.named_parameters():
        torch.testing.assert_close(fused[n], p.grad, atol=1e-4, rtol=1e-3, msg=n)


def test_add_tanh_tensor_core_large_vocab(cuda_lib):
    """V = 1024, H = 512 (BASELINE cfg 4's joint): the tcgen05 forward runs 8 vocabulary chunks with
    the online log-sum-exp across them; checked against the fp32 CUDA-core kernels (bf16 bound)."""
    d = synthetic.make_batch(2, 48, 10, 1024, 512, mode="add_tanh", ragged=True, seed=1239, device="cuda")
    a = rb.joint_rnnt_costs(d["enc"], d["dec"], d["weight"], d["bias"], d["labels"], d["act_lens"],
                            d["label_lens"], 0, "add_tanh", "bf16")
    b = rb.joint_rnnt_costs(d["enc"], d["dec"], d["weight"], d["bias"], d["labels"], d["act_lens"],
                            d["label_lens"], 0, "add_tanh", "fp32")
    torch.testing.assert_close(a, b, rtol=5e-3, atol=0)
