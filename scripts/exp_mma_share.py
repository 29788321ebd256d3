#!/usr/bin/env python
"""How much of cg_lse_mm / cg_grad_mm (concat-GELU cell kernels, cfg 2 and cfg 4) is the tensor-core product itself?

Times rnntb200_joint_cg_logprobs (factor rows + partition kernel) and rnntb200_joint_cg_bwd (gradient kernel)
through the C ABI of the product library and of an experiment build with the products compiled out
(scripts/build_exp_nomma.sh: staging, per-cell scalars, epilogues and atomics remain).  The difference bounds
what ANY faster product (an M=128 tcgen05 tiling included) could save.  L2 flushed between launches, CUDA
events.  Writes gpurun_out/<tag>_mma_share.json.
"""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rnntransducer_b200 import _lib, synthetic  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def load_exp():
    lib = ctypes.CDLL(os.path.join(ROOT, "rnntransducer_b200", "build", "exp", "librnnt_b200_nomma.so"))
    for name, (res, args) in _lib.SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2e"
    real, exp = _lib.load(), load_exp()
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    p = lambda t: t.data_ptr()
    out = {}
    for cfg in (2, 4):
        c = synthetic.CONFIGS[cfg]
        B, T, U, V, H = c["B"], c["T"], c["U"], c["V"], c["H"]
        U1 = U + 1
        d = synthetic.make_batch(B, T, U, V, H, seed=1234, device=dev)
        with torch.no_grad():
            penc = F.linear(F.gelu(d["enc"], approximate="tanh"), d["weight"][:, :H], d["bias"]).contiguous()
            pdec = F.linear(F.gelu(d["dec"], approximate="tanh"), d["weight"][:, H:]).contiguous()
        f32 = dict(device=dev, dtype=torch.float32)
        lp2, lse = torch.empty(B, T, U1, 2, **f32), torch.empty(B, T, U1, **f32)
        alpha, beta = (torch.empty(B, T, U1, device=dev, dtype=torch.int32) for _ in range(2))
        costs, gcosts = torch.empty(B, **f32), torch.full((B,), 1.0 / B, **f32)
        d_penc, d_pdec = torch.empty_like(penc), torch.empty_like(pdec)
        lab, al, ll = d["labels"], d["act_lens"], d["label_lens"]
        fac_bytes = real.rnntb200_joint_cg_factors_bytes(B, T, U1, V)
        fac = torch.empty(fac_bytes, dtype=torch.uint8, device=dev)

        def logprobs(lib, lp2_, lse_):
            _lib.check(lib.rnntb200_joint_cg_logprobs(p(penc), p(pdec), p(lab), p(al), p(ll), B, T, U1, V, 0, p(lp2_),
                                                      p(lse_), p(fac), fac_bytes, stream))

        def grad(lib):
            _lib.check(lib.rnntb200_joint_cg_bwd(p(penc), p(pdec), p(lab), p(al), p(ll), B, T, U1, V, 0, p(lse), p(alpha),
                                                 p(beta), p(gcosts), p(d_penc), p(d_pdec), 0, None, 0, p(fac), fac_bytes,
                                                 stream))

        # valid planes from the product library: the experiment gradient reads them
        logprobs(real, lp2, lse)
        _lib.check(real.rnntb200_lattice_sweep(p(lp2), p(al), p(ll), B, T, U1, p(alpha), p(beta), p(costs), None, stream))
        torch.cuda.synchronize()
        scratch_lp2, scratch_lse = torch.empty_like(lp2), torch.empty_like(lse)

        def time_us(fn, iters=30):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            tot = 0.0
            for _ in range(iters):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                e1.synchronize()
                tot += e0.elapsed_time(e1)
            return 1e3 * tot / iters

        r = {
            "logprobs_us": time_us(lambda: logprobs(real, scratch_lp2, scratch_lse)),
            "logprobs_no_product_us": time_us(lambda: logprobs(exp, scratch_lp2, scratch_lse)),
            "grad_us": time_us(lambda: grad(real)),
            "grad_no_product_us": time_us(lambda: grad(exp)),
        }
        r["logprobs_product_share"] = 1.0 - r["logprobs_no_product_us"] / r["logprobs_us"]
        r["grad_product_share"] = 1.0 - r["grad_no_product_us"] / r["grad_us"]
        out[f"cfg{cfg}"] = {k: round(v, 3) for k, v in r.items()}
        print(f"cfg{cfg}", out[f"cfg{cfg}"], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"{tag}_mma_share.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
