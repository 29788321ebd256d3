"""GPU parity of the WIDE factorised concat-GELU kernels (128 < V <= 4096, csrc/joint_cg_mm.cu): the partition
accumulates over 128-column chunks of the vocabulary, the gradient runs one CTA per (utterance, 32 frames,
128 columns).  What can go wrong that the V <= 128 tests do not see: a chunk boundary (partial last chunk,
V not a multiple of 8), labels / the blank in a chunk other than the first, several position chunks times
several column chunks in the forward's double buffer, the per-chunk slabs of the deterministic mode, tiles
entirely in the padding, the exact (underflow) path restricted to a chunk's columns.  Checked against the
fp64 CPU restatement (oracle/joint_ref.py) and against the generic per-cell kernels (RNNTB200_CG_GENERIC).
Tolerances as in test_gpu_joint_cg.py: costs 1e-5 relative, activation gradients 1e-4 absolute against fp64,
parameter gradients (sums over every lattice cell, |d_bias| reaches 50 here) conftest.param_atol = 1e-4 +
5e-5 * max|ref| (measured on these shapes: 1.5e-4 absolute = 5e-6 relative at worst).
"""
import numpy as np
import pytest
import torch

import rnntransducer_b200 as rb
from conftest import param_atol
from oracle import joint_ref
from rnntransducer_b200 import synthetic

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-5
GRAD_ATOL = 1e-4
KEYS = ("d_enc", "d_dec", "d_weight", "d_bias")


def ref_step(d, blank=0, dtype=torch.float64):
    return joint_ref.joint_loss_fwd_bwd(d["enc"], d["dec"], d["weight"], d["bias"], d["labels"].numpy(),
                                        d["act_lens"].numpy(), d["label_lens"].numpy(), blank, "mean",
                                        "concat_gelu", dtype=dtype)


def fused_step(d, blank=0, deterministic=False):
    d = {k: v.cuda() for k, v in d.items()}
    t = {k: d[k].clone().requires_grad_(True) for k in ("enc", "dec", "weight", "bias")}
    costs = rb.joint_rnnt_costs(t["enc"], t["dec"], t["weight"], t["bias"], d["labels"], d["act_lens"],
                                d["label_lens"], blank, "concat_gelu", "fp32", deterministic)
    costs.mean().backward()
    out = dict(costs=costs.detach().cpu().numpy())
    out.update({"d_" + k: v.grad.cpu().numpy() for k, v in t.items()})
    return out


def check(r, ref64, what=""):
    np.testing.assert_allclose(r["costs"], ref64["costs"], rtol=LOSS_RTOL, err_msg=what)
    for k in KEYS:
        assert np.isfinite(r[k]).all(), (what, k)
        atol = param_atol(ref64[k]) if k in ("d_weight", "d_bias") else GRAD_ATOL
        np.testing.assert_allclose(r[k], ref64[k], atol=atol, err_msg=f"{what} {k}")


# (B, T, U, V, H): V = 129 (chunks 128 + 8, one real column in the second), 136, 200, 256 (two full chunks),
# 257, 1000; U + 1 beyond 64 / 128 (several position chunks of the forward) and beyond 48 / 96 (of the backward)
SHAPES = [(2, 70, 50, 129, 16), (2, 40, 70, 256, 16), (3, 33, 9, 300, 24), (1, 45, 130, 200, 8),
          (2, 37, 20, 136, 16), (2, 21, 100, 257, 8), (2, 50, 12, 1000, 16)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("deterministic", [False, True])
def test_wide_matches_fp64_restatement(cuda_lib, oracle_lib, shape, deterministic):
    B, T, U, V, H = shape
    d = synthetic.make_batch(B, T, U, V, H, ragged=True, seed=1000 + V + U)
    if B >= 3:
        d["label_lens"][1] = 0
        d["labels"][1] = 0
        d["act_lens"][2] = 1
    assert cuda_lib.rnntb200_joint_cg_factors_bytes(B, T, U + 1, V) > 0  # the factorised kernels run
    check(fused_step(d, deterministic=deterministic), ref_step(d), str(shape))


def test_wide_labels_and_blank_in_every_chunk(cuda_lib, oracle_lib):
    """Labels drawn from the LAST columns and a blank index in the middle chunk: the one-hot corrections and
    the A[t][y] look-ups must use global column indices."""
    B, T, U, V, H = 2, 40, 30, 300, 16
    d = synthetic.make_batch(B, T, U, V, H, ragged=True, seed=4242)
    blank = 150
    g = torch.Generator().manual_seed(3)
    lab = torch.randint(0, V - 1, d["labels"].shape, generator=g, dtype=torch.int64)
    lab = lab + (lab >= blank).long()  # any column but the blank
    lab[0, : U // 2] = torch.randint(290, 300, (U // 2,), generator=g)
    lab[1, : U // 2] = torch.randint(128, 140, (U // 2,), generator=g)
    d["labels"] = lab.to(d["labels"].dtype)
    ref64 = ref_step(d, blank=blank)
    for det in (False, True):
        check(fused_step(d, blank=blank, deterministic=det), ref64, f"det={det}")


def test_wide_matches_generic_kernels(cuda_lib, monkeypatch):
    """The same inputs through the generic per-cell kernels (RNNTB200_CG_GENERIC): two independent
    evaluations of the same function on the GPU."""
    d = synthetic.make_batch(3, 64, 40, 500, 32, ragged=True, seed=99)
    wide = fused_step(d)
    monkeypatch.setenv("RNNTB200_CG_GENERIC", "1")
    assert cuda_lib.rnntb200_joint_cg_factors_bytes(3, 64, 41, 500) == 0
    generic = fused_step(d)
    np.testing.assert_allclose(wide["costs"], generic["costs"], rtol=LOSS_RTOL)
    for k in KEYS:
        np.testing.assert_allclose(wide[k], generic[k], err_msg=k,
                                   atol=param_atol(generic[k]) if k in ("d_weight", "d_bias") else GRAD_ATOL)


@pytest.mark.parametrize("scale", [8.0, 40.0])
def test_wide_extreme_logit_ranges(cuda_lib, oracle_lib, scale):
    """Row peaks that do not line up: cells whose factorised partition underflows take the exact path, which
    in the wide gradient is restricted to each CTA's own columns."""
    d = synthetic.make_batch(3, 21, 7, 150, 16, ragged=True, seed=321)
    d["weight"] = d["weight"] * scale
    ref64 = ref_step(d)
    for det in (False, True):
        r = fused_step(d, deterministic=det)
        np.testing.assert_allclose(r["costs"], ref64["costs"], rtol=2e-5)
        for k in KEYS:
            np.testing.assert_allclose(r[k], ref64[k], atol=param_atol(ref64[k], 1e-4 * scale), err_msg=k)


def test_wide_deterministic_mode_is_bit_reproducible(cuda_lib):
    d = synthetic.make_batch(3, 64, 17, 400, 32, ragged=True, seed=5)
    a, b = fused_step(d, deterministic=True), fused_step(d, deterministic=True)
    for k in ("costs",) + KEYS:
        assert np.array_equal(a[k], b[k]), k


def test_very_large_vocabulary(cuda_lib, oracle_lib):
    """V = 4100 (33 column chunks, the last one 8 wide with 4 real columns): there is no upper limit."""
    B, T, U, V, H = 1, 10, 3, 4100, 8
    d = synthetic.make_batch(B, T, U, V, H, ragged=False, seed=8)
    assert cuda_lib.rnntb200_joint_cg_factors_bytes(B, T, U + 1, V) > 0
    check(fused_step(d), ref_step(d), "V=4100")
