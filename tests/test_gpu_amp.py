"""GPU: the fused path under the reference's shipped precision (scripts/run_train.sh:32 --precision=16 ->
model.py:28-31: fp16 autocast, torchaudio's loss).

The projection kernels take fp16 / bf16 encoder / predictor outputs directly (converted exactly on load,
gradients rounded to nearest on store, fp32 arithmetic in between), so:
  * results on half inputs are those of the fp32 path fed the same values up-cast (costs, d_enc, d_dec bit for
    bit; d_weight / d_bias to fp32 atomic-accumulation order) -- the AMP mode inherits the fp32 path's parity,
    there is no separate half arithmetic to validate;
  * under torch.autocast the drop-in module trains like the reference's --precision=16 configuration, whose
    own GPU path (eager joint under autocast + torchaudio CUDA rnnt_loss on fp16 logits) is the like-for-like
    comparison below (tolerance: fp16 rounding of the reference's logits, not of ours).
"""
import numpy as np
import pytest
import torch

import rnntransducer_b200 as rb
from rnntransducer_b200 import synthetic

pytestmark = pytest.mark.gpu


def _step(d, enc, dec):
    t = dict(enc=enc.clone().requires_grad_(True), dec=dec.clone().requires_grad_(True),
             weight=d["weight"].clone().requires_grad_(True), bias=d["bias"].clone().requires_grad_(True))
    costs = rb.joint_rnnt_costs(t["enc"], t["dec"], t["weight"], t["bias"], d["labels"], d["act_lens"], d["label_lens"],
                                0, "concat_gelu", "fp32", True)
    costs.mean().backward()
    return costs.detach(), {k: v.grad for k, v in t.items()}


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(3, 50, 9, 73, 128), (4, 130, 20, 73, 512)])
def test_half_activations_are_the_fp32_path_on_the_same_values(cuda_lib, dtype, shape):
    B, T, U, V, H = shape
    d = synthetic.make_batch(B, T, U, V, H, ragged=True, seed=31, device="cuda")
    enc_h, dec_h = d["enc"].to(dtype), d["dec"].to(dtype)
    c_h, g_h = _step(d, enc_h, dec_h)
    c_f, g_f = _step(d, enc_h.float(), dec_h.float())
    assert g_h["enc"].dtype == dtype and g_h["dec"].dtype == dtype and g_h["weight"].dtype == torch.float32
    assert torch.equal(c_h, c_f)
    assert torch.equal(g_h["enc"], g_f["enc"].to(dtype)) and torch.equal(g_h["dec"], g_f["dec"].to(dtype))
    # d_weight / d_bias: the same products, but summed across tiles with fp32 atomics (order varies run to run)
    for k in ("weight", "bias"):
        torch.testing.assert_close(g_h[k], g_f[k], rtol=0, atol=2e-6 * float(g_f[k].abs().max()) + 1e-7)


def test_autocast_module_against_the_references_fp16_gpu_path(cuda_lib):
    torchaudio = pytest.importorskip("torchaudio")
    torch.manual_seed(2)
    ep = dict(input_size=16, hidden_size=32, output_size=128, num_layers=1, rnn_type="lstm", dropout=0.0, bidirectional=True)
    dp = dict(embedding_size=29, pad_token_id=0, hidden_size=32, output_size=128, num_layers=1, rnn_type="lstm", dropout=0.0)
    net = rb.JointNet(ep, dp, 29).cuda().train()
    from rnntransducer_b200.training import synthetic_training_batch
    audios, al, tal, texts, tl, targets, tgl = synthetic_training_batch(4, 60, 12, 16, 29, ragged=True, seed=3)
    audios, tal, texts, targets, tgl = (x.cuda() for x in (audios, tal, texts, targets, tgl))
    loss_fn = rb.RNNTLoss(0, "mean", warp_compat=False)
    with torch.autocast("cuda", dtype=torch.float16):
        logits = net(audios, al, texts, tl)
        assert isinstance(logits, rb.JointLogits) and logits.enc.dtype == torch.float16
        loss = loss_fn(logits, targets, tal, tgl)
    loss.backward()
    ours = {n: p.grad.clone() for n, p in net.named_parameters()}
    net.zero_grad()
    net.fused = False
    with torch.autocast("cuda", dtype=torch.float16):
        dense = net(audios, al, texts, tl)  # the reference's own joint under autocast: fp16 logits
    ref = torchaudio.functional.rnnt_loss(dense, targets, tal, tgl, blank=0, reduction="mean")
    ref.backward()
    assert abs(float(loss) - float(ref)) < 5e-3 * abs(float(ref))
    for n, p in net.named_parameters():
        scale = max(float(p.grad.abs().max()), 1e-3)
        assert float((ours[n] - p.grad).abs().max()) < 3e-2 * scale, n
