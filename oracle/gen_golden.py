"""Generates tests/golden/*.npz from the REAL reference + torchaudio, in the build container.

Run:  python -m oracle.gen_golden         (needs /root/reference; NOT run on the GPU box)

What pins what (the reference has no tests of its own for this path):
  kat1.npz          classic warp-transducer / torchaudio docstring example
                    (torchaudio/transforms/_transforms.py:1812-1826), cost + grad recomputed here
                    with torchaudio CPU rnnt_loss (the loss of model.py:6,31).
  dense_*.npz       random dense logits, full and ragged lengths (incl. U_b = 0 and T_b = 1):
                    torchaudio CPU costs + grads.
  joint_*.npz       the reference's own JointNet.joint (imported from /root/reference/networks,
                    transducer.py:41-71, with a 3-symbol pyctcdecode stub) on random enc/dec and
                    random fc parameters -> logits; torchaudio CPU loss on those logits;
                    autograd grads w.r.t. enc, dec, fc.weight, fc.bias.
  addtanh_*.npz     torchaudio.models.rnnt._Joiner(activation="tanh") + torchaudio loss
                    (oracle of the add_tanh mode; not a reference function).
  jointnet_fwd.npz  full reference JointNet.forward (GRU encoder + LSTM predictor) with a fixed
                    state_dict: checks that the drop-in JointNet loads reference checkpoints and
                    reproduces logits (transducer.py:73-93).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch
import torchaudio

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from rnntransducer_b200 import synthetic  # noqa: E402


def import_reference_jointnet():
    for name, attrs in (("pyctcdecode", dict(LanguageModel=object)),
                        ("pyctcdecode.language_model", dict(HotwordScorer=object)),
                        ("pyctcdecode.constants", dict(DEFAULT_HOTWORD_WEIGHT=10.0))):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules.setdefault(name, m)
    sys.path.insert(0, "/root/reference")
    from networks import JointNet  # the reference's own class
    return JointNet


def ta_loss(logits, labels, act_lens, label_lens, blank=0):
    """torchaudio CPU rnnt_loss: per-utterance costs + grad of sum(costs) wrt logits."""
    x = logits.detach().clone().requires_grad_(True)
    c = torchaudio.functional.rnnt_loss(x, labels, act_lens, label_lens, blank=blank,
                                        reduction="none")
    c.sum().backward()
    return c.detach(), x.grad


def save(name, **arrays):
    path = os.path.join(GOLD, name)
    np.savez_compressed(path, **{k: (v.numpy() if torch.is_tensor(v) else np.asarray(v))
                                 for k, v in arrays.items()})
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(4)

    # --- KAT-1 -------------------------------------------------------------------------------
    logits = torch.tensor([[[[.1, .6, .1, .1, .1], [.1, .1, .6, .1, .1], [.1, .1, .2, .8, .1]],
                            [[.1, .6, .1, .1, .1], [.1, .1, .2, .1, .1], [.7, .1, .2, .1, .1]]]])
    labels = torch.tensor([[1, 2]], dtype=torch.int32)
    lens = torch.tensor([2], dtype=torch.int32)
    c, g = ta_loss(logits, labels, lens, lens)
    assert abs(float(c[0]) - 4.49566698) < 1e-6
    save("kat1.npz", logits=logits, labels=labels, act_lens=lens, label_lens=lens, costs=c, grads=g,
         blank=0)

    # --- dense logits ------------------------------------------------------------------------
    cases = {
        "dense_full": dict(B=3, T=9, U=4, V=7, ragged=False, seed=11, blank=0),
        "dense_ragged": dict(B=5, T=13, U=6, V=11, ragged=True, seed=12, blank=0),
        "dense_blank_last": dict(B=2, T=7, U=3, V=6, ragged=True, seed=13, blank=5),
        "dense_v73": dict(B=2, T=20, U=8, V=73, ragged=True, seed=14, blank=0),
    }
    for name, cfg in cases.items():
        blank = cfg.pop("blank")
        d = synthetic.make_dense_logits(**cfg)
        if blank != 0:  # labels must avoid the blank id
            d["labels"] = (d["labels"] - 1) % (cfg["V"] - 1)
        if name == "dense_ragged":  # the edge cases the oracle was probed on
            d["label_lens"][1] = 0
            d["act_lens"][2] = 1
            d["label_lens"][3] = cfg["U"]
            d["act_lens"][3] = 1
        c, g = ta_loss(d["logits"], d["labels"], d["act_lens"], d["label_lens"], blank)
        save(name + ".npz", **d, costs=c, grads=g, blank=blank)

    # --- reference JointNet.joint + loss ------------------------------------------------------
    JointNet = import_reference_jointnet()
    enc_p = dict(input_size=8, hidden_size=8, output_size=16, num_layers=1, rnn_type="gru",
                 dropout=0.0, bidirectional=True)
    dec_p = dict(embedding_size=11, pad_token_id=0, hidden_size=8, output_size=16, num_layers=1,
                 rnn_type="lstm", dropout=0.0)

    def joint_case(name, B, T, U, V, H, ragged, seed, subsample=None):
        d = synthetic.make_batch(B, T, U, V, H, mode="concat_gelu", ragged=ragged, seed=seed)
        ep, dp = dict(enc_p, output_size=H), dict(dec_p, output_size=H, embedding_size=V)
        net = JointNet(ep, dp, V)
        with torch.no_grad():
            net.fc.weight.copy_(d["weight"])
            net.fc.bias.copy_(d["bias"])
        enc = d["enc"].clone().requires_grad_(True)
        dec = d["dec"].clone().requires_grad_(True)
        logits = net.joint(enc, dec)  # transducer.py:41-71, the reference's own code
        costs = torchaudio.functional.rnnt_loss(logits, d["labels"], d["act_lens"], d["label_lens"],
                                                blank=0, reduction="none")
        loss = costs.mean()  # reduction="mean" of model.py:31,39
        loss.backward()
        out = dict(costs=costs.detach(), loss=loss.detach(), d_enc=enc.grad, d_dec=dec.grad,
                   d_weight=net.fc.weight.grad, d_bias=net.fc.bias.grad)
        if subsample is None:
            save(name, **d, logits=logits.detach(), **out)
        else:  # big case: inputs are regenerated from the seed; keep outputs sub-sampled
            st, sh = subsample
            save(name, seed=seed, B=B, T=T, U=U, V=V, H=H, ragged=ragged, costs=out["costs"],
                 loss=out["loss"], d_bias=out["d_bias"], d_weight=out["d_weight"][:, ::sh],
                 d_enc=out["d_enc"][:, ::st, ::sh], d_dec=out["d_dec"][:, :, ::sh],
                 stride_t=st, stride_h=sh, enc_sum=d["enc"].double().sum(),
                 logits_sub=logits.detach()[:, ::st, ::5, :])

    joint_case("joint_small_full.npz", B=2, T=12, U=5, V=11, H=16, ragged=False, seed=21)
    joint_case("joint_small_ragged.npz", B=4, T=17, U=6, V=11, H=16, ragged=True, seed=22)
    c1 = synthetic.CONFIGS[1]
    joint_case("joint_cfg1.npz", c1["B"], c1["T"], c1["U"], c1["V"], c1["H"], ragged=False,
               seed=1235, subsample=(10, 8))
    joint_case("joint_cfg1_ragged.npz", c1["B"], c1["T"], c1["U"], c1["V"], c1["H"], ragged=True,
               seed=1235, subsample=(10, 8))

    # --- add_tanh: torchaudio _Joiner + loss --------------------------------------------------
    from torchaudio.models.rnnt import _Joiner

    def addtanh_case(name, B, T, U, V, H, ragged, seed):
        d = synthetic.make_batch(B, T, U, V, H, mode="add_tanh", ragged=ragged, seed=seed)
        joiner = _Joiner(H, V, activation="tanh")
        with torch.no_grad():
            joiner.linear.weight.copy_(d["weight"])
            joiner.linear.bias.copy_(d["bias"])
        enc = d["enc"].clone().requires_grad_(True)
        dec = d["dec"].clone().requires_grad_(True)
        logits, _, _ = joiner(enc, d["act_lens"], dec, d["label_lens"])
        costs = torchaudio.functional.rnnt_loss(logits, d["labels"], d["act_lens"], d["label_lens"],
                                                blank=0, reduction="none")
        loss = costs.mean()
        loss.backward()
        save(name, **d, logits=logits.detach(), costs=costs.detach(), loss=loss.detach(),
             d_enc=enc.grad, d_dec=dec.grad, d_weight=joiner.linear.weight.grad,
             d_bias=joiner.linear.bias.grad)

    addtanh_case("addtanh_small_full.npz", B=2, T=12, U=5, V=11, H=16, ragged=False, seed=31)
    addtanh_case("addtanh_small_ragged.npz", B=3, T=15, U=6, V=13, H=32, ragged=True, seed=32)

    # --- full reference JointNet.forward with a fixed state_dict --------------------------------
    torch.manual_seed(41)
    ep = dict(input_size=8, hidden_size=12, output_size=16, num_layers=2, rnn_type="gru",
              dropout=0.0, bidirectional=True)
    dp = dict(embedding_size=11, pad_token_id=0, hidden_size=12, output_size=16, num_layers=2,
              rnn_type="lstm", dropout=0.0)
    net = JointNet(dict(ep), dict(dp), 11).eval()
    audio = torch.randn(3, 14, 8)
    audio_lengths = [14, 9, 11]
    texts = torch.tensor([[0, 3, 4, 5, 1], [0, 2, 2, 0, 0], [0, 7, 8, 9, 0]], dtype=torch.int64)
    text_lengths = [5, 3, 4]
    with torch.no_grad():
        logits = net(audio, audio_lengths, texts, text_lengths)
        one_d = net.joint(torch.ones(16) * 0.3, torch.ones(16) * -0.2)  # decode form, :125
    sd = {"sd__" + k: v for k, v in net.state_dict().items()}
    save("jointnet_fwd.npz", audio=audio, audio_lengths=np.array(audio_lengths), texts=texts,
         text_lengths=np.array(text_lengths), logits=logits, joint_1d=one_d, **sd)


if __name__ == "__main__":
    main()
