#!/usr/bin/env python
"""One small forward+backward through every kernel family, for compute-sanitizer:
    compute-sanitizer --tool memcheck python scripts/sanitize_small.py
Shapes are chosen to hit ragged borders, partial tiles, multi-warp / cluster sweeps, multi-chunk
vocabularies and both tensor-core joints."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rnntransducer_b200 as rb  # noqa: E402
from rnntransducer_b200 import synthetic  # noqa: E402


def step(d, mode, gemm="fp32", det=False):
    t = {k: d[k].clone().requires_grad_(True) for k in ("enc", "dec", "weight", "bias")}
    loss = rb.joint_rnnt_loss(t["enc"], t["dec"], t["weight"], t["bias"], d["labels"], d["act_lens"],
                              d["label_lens"], 0, "mean", mode, gemm, deterministic=det)
    loss.backward()
    torch.cuda.synchronize()
    return float(loss)


def main():
    dev = "cuda"
    # dense loss: single warp, multi warp, cluster sweep
    for (B, T, U, V) in ((3, 17, 5, 11), (2, 40, 70, 9), (2, 30, 200, 6)):
        d = synthetic.make_dense_logits(B, T, U, V, ragged=True, seed=1, device=dev)
        x = d["logits"].requires_grad_(True)
        rb.rnnt_loss(x, d["labels"], d["act_lens"], d["label_lens"], 0, "mean").backward()
        torch.cuda.synchronize()
        print("dense", (B, T, U, V), "ok")
    # reference-exact joint: factorised kernels + tensor-core projections, generic kernels (V > 128)
    for (B, T, U, V, H) in ((3, 45, 9, 73, 128), (2, 37, 50, 73, 64), (2, 20, 6, 200, 32)):
        d = synthetic.make_batch(B, T, U, V, H, ragged=True, seed=2, device=dev)
        for det in (False, True):
            print("concat_gelu", (B, T, U, V, H), det, step(d, "concat_gelu", det=det))
    # add-tanh joint: CUDA-core, tcgen05 fwd + bwd, multi-chunk vocabulary
    for (B, T, U, V, H, gemm) in ((2, 21, 9, 73, 48, "fp32"), (2, 33, 9, 73, 128, "bf16"), (2, 18, 10, 200, 128, "bf16"),
                                 (1, 20, 5, 73, 64, "bf16")):
        d = synthetic.make_batch(B, T, U, V, H, mode="add_tanh", ragged=True, seed=3, device=dev)
        print("add_tanh", (B, T, U, V, H), gemm, step(d, "add_tanh", gemm))
    print("sanitize_small OK")


if __name__ == "__main__":
    main()
