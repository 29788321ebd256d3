#!/usr/bin/env python
"""bench.py -- fused joint + RNN-T loss, forward + backward, in lattice cells/s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--cfg 2] [--mode concat_gelu|add_tanh] [--gemm fp32|bf16] [--ragged] [--eager]

One "step" = one pass of the hot path over one synthetic batch: joint + log-softmax + alpha/beta
sweeps + gradient w.r.t. enc, dec, fc.weight, fc.bias (reduction="mean"), i.e. what
``loss = RNNTLoss(...)(JointNet(...)(...), ...); loss.backward()`` costs from the encoder /
predictor outputs down (reference model.py:56-57, networks/transducer.py:54-71).

Default workload at N=1: BASELINE.json configs[1] -- KsponSpeech-shaped batch B=32, T=400, U=80,
V=73, joint H=512, fp32, full-length utterances, the reference's own joint (concat -> GELU ->
Linear).  Under torchrun every rank runs that batch with its own seed (weak scaling; utterances
are independent) and the only exchange is the all-reduce of the fc gradients DDP would do.

Printed keys (one JSON line on rank 0): the base contract plus
  roofline      dominant kernel, algorithmic bytes per launch / CUDA-event duration vs measured HBM peak
  kernels       the same for every kernel of the step
  cpu_baseline  the CPU oracle port (reference joint restated in torch + C/OpenMP warp-transducer
                restatement) timed on this box's host cores on a bounded sample of the workload
  e2e           same metric through the public API with HOST (pinned) inputs, H2D + D2H in the timing
``--impl reference`` times that CPU path alone (rank 0 only) and prints the same line shape.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "joint+RNNT-loss fwd+bwd lattice cells/s"
UNIT = "cells/s"
FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=100)
    p.add_argument("--warmup", type=int, default=10)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--cfg", type=int, default=2, choices=[1, 2, 3, 4, 5],
                   help="BASELINE.json configs[cfg-1]; 5 = full training step (LSTM encoder/predictor + fused joint/loss, DDP)")
    p.add_argument("--mode", default="concat_gelu", choices=["concat_gelu", "add_tanh"])
    p.add_argument("--gemm", default=None, choices=["fp32", "bf16"])
    p.add_argument("--ragged", action="store_true")
    p.add_argument("--batch", type=int, default=None, help="override the config's utterances per GPU")
    p.add_argument("--eager", action="store_true", help="do not replay the step from a CUDA graph")
    p.add_argument("--deterministic", action="store_true")
    p.add_argument("--act-dtype", default="fp32", choices=["fp32", "fp16", "bf16"],
                   help="dtype of the encoder / predictor outputs handed to the joint (fp16 = the reference's shipped "
                        "--precision=16, scripts/run_train.sh:32): concat_gelu reads them directly, H2D bytes halve")
    p.add_argument("--allreduce", default="peer", choices=["peer", "nccl"],
                   help="N>1: gradient all-reduce of the fc grads -- one-shot kernel over NVLink peer memory inside the "
                        "step's CUDA graph (default) or torch.distributed/NCCL issued from the host after the graph")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-gpu-baseline", action="store_true",
                   help="skip the reference's own GPU path (eager joint + torchaudio CUDA rnnt_loss)")
    p.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU work for cpu_baseline")
    return p.parse_args()


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, c, args):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture
    (profiles/dram_traffic.json, written by scripts/summarize_profiles.py).  The capture is of the
    default workload (cfg 2, full-length): null for anything else."""
    if args.cfg != 2 or args.batch or args.ragged:
        return None
    keys = {"proj_tc_kernel": "proj_tc_kernel", "proj_tc_bwd_kernel": "proj_tc_bwd_kernel", "cg_lse_kernel": "cg_lse", "cg_grad_kernel": "cg_grad",
            "lattice_sweep_kernel": "lattice_sweep", "at_lse_kernel": "at_lse_tc", "at_grad_kernel": "at_grad_tc"}
    try:
        with open(os.path.join(ROOT, "profiles", "dram_traffic.json")) as f:
            return json.load(f)[keys[kernel]]["dram_bytes_per_launch"]
    except Exception:
        return None


def workload_config(args, world):
    from rnntransducer_b200 import synthetic
    c = dict(synthetic.CONFIGS[2 if args.cfg == 5 else args.cfg])  # cfg 5: the cfg-2 batch per rank, full model
    if args.batch:
        c["B"] = args.batch
    gemm = args.gemm or ("fp32" if args.mode == "concat_gelu" or args.cfg == 2 else "bf16")
    return c, gemm, {
        "workload": f"BASELINE cfg{args.cfg}: B={c['B']} T={c['T']} U={c['U']} V={c['V']} H={c['H']} per GPU, "
                    f"{'ragged' if args.ragged else 'full-length'} utterances, fwd+bwd",
        "joint": args.mode, "gemm": gemm, "B_per_gpu": c["B"], "global_batch": c["B"] * world,
        "T": c["T"], "U": c["U"], "V": c["V"], "H": c["H"], "ragged": bool(args.ragged),
        "reduction": "mean", "parallelism": f"dp{world} (utterances sharded, fc grads all-reduced)",
        "l2": "flushed between timed steps (256 MiB write)",
    }


# ------------------------------------------------------------------------------------------------
# clocks
class ClockSampler:
    """Samples SM clock + throttle reasons of one GPU during the timed region (pynvml)."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index, period=0.01):
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._active = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self._sample()  # first-use costs of the queries are paid here, not inside the timed region
            self.samples.clear()
            self.reasons.clear()
        except Exception:
            self.nv = None

    def _sample(self):
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            try:
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            if self._active.is_set():
                self._sample()
                self._stop.wait(self.period)
            else:
                self._active.wait(0.05)

    def start(self):
        """Start the sampling thread (idle until the timed region is entered).  Called BEFORE the barrier that
        precedes the timed region, like everything else whose duration differs from rank to rank."""
        if self.nv is not None and self._thr is None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __enter__(self):
        self.start()
        self._active.set()
        return self

    def __exit__(self, *exc):
        if self._thr is not None:
            self._sample()
            self._active.clear()
            self._stop.set()
            self._active.set()  # wake the thread so that it sees the stop flag
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unsampled"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------
# CPU arm (the oracle port: reference joint restated + warp-transducer CPU restatement)
def cpu_step_factory(batch, n_utt, mode):
    """Returns (fn, cells) where fn() runs joint + loss + backward on the first n_utt utterances."""
    import torch
    from oracle import joint_ref
    from rnntransducer_b200 import synthetic
    sl = slice(0, n_utt)
    enc, dec = batch["enc"][sl].contiguous(), batch["dec"][sl].contiguous()
    labels = batch["labels"][sl].contiguous().numpy()
    al, ll = batch["act_lens"][sl].contiguous().numpy(), batch["label_lens"][sl].contiguous().numpy()
    cells = synthetic.count_cells(batch["act_lens"][sl], batch["label_lens"][sl])

    def fn():
        return joint_ref.joint_loss_fwd_bwd(enc, dec, batch["weight"], batch["bias"], labels, al, ll, 0,
                                            "mean", mode, num_threads=0)
    return fn, cells


def cpu_plan(batch, mode, budget_seconds, n_calls):
    """How many utterances one CPU step takes so that `n_calls` steps fit `budget_seconds`.

    One untimed call first (builds / loads the oracle library, spins up the OpenMP and torch thread
    pools, faults the pages in), then ONE timed call on min(cores, B) utterances -- the loss
    parallelises over utterances, so anything smaller leaves cores idle -- and from its per-utterance
    time the largest multiple of the thread count (up to the whole batch) that fits the budget."""
    import torch
    from oracle import warp_cpu
    warp_cpu.build()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = batch["enc"].shape[0]
    threads = min(cores, B)
    cpu_step_factory(batch, 1, mode)[0]()          # warm-up, untimed
    fn, _ = cpu_step_factory(batch, threads, mode)
    t0 = time.perf_counter()
    fn()
    t_round = time.perf_counter() - t0            # one utterance per thread
    rounds = int(max(1, min(B // threads, budget_seconds / max(n_calls, 1) / max(t_round, 1e-3))))
    n = rounds * threads if rounds * threads < B else B
    return n, threads, cores


def time_torchaudio_cpu(batch, mode, n_utt=2):
    """torchaudio's CPU rnnt_loss (the loss of the reference's shipped fp16 path, model.py:6,31; BASELINE.md
    section 4 item 2) behind the same restated joint, fwd+bwd, once, on `n_utt` utterances."""
    import torch
    try:
        import torchaudio
    except Exception as e:  # pragma: no cover
        return {"unavailable": repr(e)}
    from oracle import joint_ref
    from rnntransducer_b200 import synthetic
    sl = slice(0, n_utt)
    leaves = [batch[k][sl].clone().requires_grad_(True) for k in ("enc", "dec")]
    w, b = batch["weight"].clone().requires_grad_(True), batch["bias"].clone().requires_grad_(True)
    cells = synthetic.count_cells(batch["act_lens"][sl], batch["label_lens"][sl])
    t0 = time.perf_counter()
    logits = joint_ref.JOINTS[mode](leaves[0], leaves[1], w, b)
    loss = torchaudio.functional.rnnt_loss(logits, batch["labels"][sl].contiguous(), batch["act_lens"][sl].contiguous(),
                                           batch["label_lens"][sl].contiguous(), blank=0, reduction="mean")
    loss.backward()
    dt = time.perf_counter() - t0
    return {"value": cells / dt, "unit": UNIT, "seconds": dt, "utterances": n_utt,
            "what": "restated joint (torch CPU) + torchaudio.functional.rnnt_loss on CPU, fwd+bwd, one run"}


def time_cpu(batch, mode, target_seconds, reps=2):
    """Bounded sample of the workload batch on this box's host cores (see cpu_plan)."""
    n, threads, cores = cpu_plan(batch, mode, target_seconds, reps)
    B = batch["enc"].shape[0]
    fn, cells = cpu_step_factory(batch, n, mode)
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return dict(value=cells / best, unit=UNIT, cores=cores, kind="port",
                sample=f"{n} of {B} utterances of the same batch ({cells} cells), joint restated in "
                       f"torch CPU ({cores} threads) + oracle/warp_cpu.c OpenMP over utterances "
                       f"({threads} threads), fwd+bwd, best of {reps} after one untimed warm-up call",
                seconds=best, utterances=n, threads=threads,
                torchaudio_cpu=time_torchaudio_cpu(batch, mode))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from rnntransducer_b200 import synthetic
    c, gemm, config = workload_config(args, world)
    batch = synthetic.make_batch(c["B"], c["T"], c["U"], c["V"], c["H"], mode=args.mode,
                                 ragged=args.ragged, seed=1234 + args.cfg)
    # the whole batch per step when (steps + warmup) of those end within ~4 minutes, else the largest
    # multiple of the thread count that does
    n, threads, cores = cpu_plan(batch, args.mode, 240.0, args.steps + args.warmup)
    fn, cells = cpu_step_factory(batch, n, args.mode)
    for _ in range(args.warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    total = time.perf_counter() - t0
    value = cells * args.steps / total
    sample = (f"each step = {n} of {c['B']} utterances of the workload batch ({cells} cells): reference "
              f"joint restated in torch CPU ({cores} threads) + oracle/warp_cpu.c (C/OpenMP restatement of "
              f"warp-transducer's CPU loss, {threads} threads over utterances), fwd+bwd")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config,
        "utterances_per_s": n * args.steps / total,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "threads": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "warprnnt_pytorch (the reference's fp32 loss) is not installable offline and the reference "
                "is pure Python, so the reference arm is the CPU oracle port (DESIGN.md); the reference's GPU "
                "path (eager joint + torchaudio CUDA rnnt_loss) is timed by the default arm as gpu_baseline",
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    import rnntransducer_b200 as rb
    from rnntransducer_b200 import _lib, synthetic
    from rnntransducer_b200.joint_add_tanh import GEMMS

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl ours) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()  # fails loudly when the CUDA extension is missing

    c, gemm, config = workload_config(args, world)
    mode, det = args.mode, bool(args.deterministic)
    B, T, U, V, H = c["B"], c["T"], c["U"], c["V"], c["H"]
    U1 = U + 1
    host = synthetic.make_batch(B, T, U, V, H, mode=mode, ragged=args.ragged, seed=1234 + args.cfg + rank)
    host32 = dict(host)  # fp32 copy for the CPU legs
    act_dtype = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}[args.act_dtype]
    if act_dtype != torch.float32:
        if mode != "concat_gelu":
            raise SystemExit("--act-dtype: only the concat_gelu joint reads half activations directly")
        host["enc"], host["dec"] = host["enc"].to(act_dtype), host["dec"].to(act_dtype)
        host32["enc"], host32["dec"] = host["enc"].float(), host["dec"].float()  # the same values
        config["activations"] = args.act_dtype
    cells = synthetic.count_cells(host["act_lens"], host["label_lens"])
    # Per-step inputs live in ONE slab (256-byte aligned fields): one pinned host slab, one device slab
    # whose views are the static buffers the CUDA graph is captured on -> the end-to-end leg uploads a
    # step's inputs with ONE host->device copy.
    per_step = ("enc", "dec", "labels", "act_lens", "label_lens")
    offs, total = {}, 0
    for k in per_step:
        offs[k] = total
        total += (host[k].numel() * host[k].element_size() + 255) // 256 * 256
    slab_host = torch.empty(total, dtype=torch.uint8).pin_memory()
    slab_dev = torch.empty(total, dtype=torch.uint8, device=dev)

    def views(slab):
        return {k: slab[offs[k]:offs[k] + host[k].numel() * host[k].element_size()].view(host[k].dtype).view(host[k].shape)
                for k in per_step}

    pinned = views(slab_host)
    for k in per_step:
        pinned[k].copy_(host[k])
    slab_dev.copy_(slab_host)
    st = {k: v.detach() for k, v in views(slab_dev).items()}  # static device buffers (graph inputs)
    st["weight"], st["bias"] = host["weight"].to(dev), host["bias"].to(dev)
    for k in ("enc", "dec", "weight", "bias"):
        st[k].requires_grad_(True)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def zero_grads():
        for k in ("enc", "dec", "weight", "bias"):
            st[k].grad = None

    out = {}
    # N > 1: what DDP does with the gradients our kernels produce (fc.weight, fc.bias) -- averaged over the
    # ranks every step, inside the timed bracket.  Default: ONE kernel over NVLink peer memory (csrc/comm.cu)
    # that is part of the captured step, so a step stays one graph replay with no host work in between.
    # --allreduce nccl: one flat bucket + one torch.distributed all_reduce issued after the graph (round 1).
    peer_ar, ar_note = None, None
    if world > 1 and args.allreduce == "peer":
        try:
            from rnntransducer_b200.comm import PeerAllReduce
            peer_ar = PeerAllReduce(st["weight"].numel() + st["bias"].numel())
        except Exception as e:  # cudaIpc not permitted on this box: say so and use NCCL
            ar_note = f"peer-memory all-reduce unavailable ({e!r}); NCCL used"
        ok = torch.tensor([int(peer_ar is not None)], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok) == 0:
            peer_ar = None

    seed = torch.ones(1, device=dev)

    def step():
        loss = rb.joint_rnnt_loss(st["enc"], st["dec"], st["weight"], st["bias"], st["labels"],
                                  st["act_lens"], st["label_lens"], 0, "mean", mode, gemm,
                                  deterministic=det)
        loss.backward(seed)  # (a cached ones tensor: without it autograd fills a fresh one every step -- one more launch)
        if peer_ar is not None:
            peer_ar.all_reduce_mean_([st["weight"].grad, st["bias"].grad])
        out["loss"] = loss.detach()

    # eager warm-up (also JITs nothing: the library is prebuilt) on a side stream, then capture
    graph = None
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            zero_grads()
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    if not args.eager:
        zero_grads()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()

    bucket = torch.empty(st["weight"].numel() + st["bias"].numel(), device=dev)

    def run_step():
        if graph is not None:
            graph.replay()
        else:
            zero_grads()
            step()
        if world > 1 and peer_ar is None:
            torch.cat((st["weight"].grad.reshape(-1), st["bias"].grad), out=bucket)
            dist.all_reduce(bucket)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev_pool = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]

    def timed(fn, n):
        ev = ev_pool[:n]
        for e0, e1 in ev:
            flush_buf.zero_()  # evict the working set from L2 (outside the timed bracket)
            e0.record()
            fn()
            e1.record()
        torch.cuda.synchronize()
        per_step.clear()
        per_step.extend(e0.elapsed_time(e1) for e0, e1 in ev)
        return sum(per_step)  # ms

    per_step = []
    # NVML is initialised and the timing events are created BEFORE the barrier: whatever a rank does between
    # the barrier and its first timed step is rank skew that every other rank waits for in step 1's collective
    # (nvmlInit on an 8-GPU box takes milliseconds and a different number of them on every rank: round 1's
    # N = 8 lines each carried one ~3 ms first step, 0.36 instead of 0.21 ms per step over 20 steps)
    clocks = ClockSampler(local_rank).start()
    for _ in range(max(args.warmup, 3)):
        flush_buf.zero_()
        run_step()
    barrier()
    with clocks:
        total_ms = timed(run_step, args.steps)
    barrier()
    step_times = sorted(per_step)

    # N > 1: this rank's step WITHOUT the collective (its own graph, nobody to wait for) -- with the
    # collective alone (below) it separates what a step costs from what waiting for the slowest rank costs
    compute_only_ms = None
    if world > 1:
        saved_ar, peer_ar = peer_ar, None
        saved_loss, saved_grads = out["loss"], {k: st[k].grad for k in ("enc", "dec", "weight", "bias")}
        zero_grads()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        zero_grads()
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g2):
            step()
        peer_ar = saved_ar
        out["loss"] = saved_loss  # the timed graph's own outputs again (the end-to-end leg reads them)
        for k, g in saved_grads.items():
            st[k].grad = g
        for _ in range(5):
            flush_buf.zero_()
            g2.replay()
        compute_only_ms = timed(g2.replay, args.steps) / args.steps
        step_times_local = sorted(per_step)
        barrier()

    # ---- end to end through the public API with host inputs ------------------------------------
    # Every step uploads its own inputs from pinned host memory and the host reads the step's loss.
    # The upload of step i+1 runs on a copy stream while step i computes (double-buffered staging,
    # then a device-side copy into the buffers the CUDA graph was captured on), as a training input
    # pipeline would; the timed region covers all K uploads, K steps and K loss reads.
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()
    h2d = slab_host.numel()
    copy_stream = torch.cuda.Stream()
    staging = [torch.empty_like(slab_dev) for _ in range(2)]
    uploaded = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def upload(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])  # staging buffer free again
            staging[i % 2].copy_(slab_host, non_blocking=True)  # the step's inputs: one H2D copy
            uploaded[i % 2].record(copy_stream)

    def e2e_run(n_steps):
        main = torch.cuda.current_stream()
        for ev in consumed:
            ev.record(main)
        upload(0)
        for i in range(n_steps):
            if i + 1 < n_steps:
                upload(i + 1)
            main.wait_event(uploaded[i % 2])
            with torch.no_grad():
                slab_dev.copy_(staging[i % 2], non_blocking=True)  # into the buffers the graph reads
            consumed[i % 2].record(main)
            run_step()
            loss_host.copy_(out["loss"].reshape(1), non_blocking=True)
            main.synchronize()  # the caller reads this step's loss
        return float(loss_host[0])

    e2e_run(3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.perf_counter()
    e0.record()
    e2e_run(args.steps)
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = e0.elapsed_time(e1)
    e2e_wall = time.perf_counter() - t_wall
    barrier()
    loss_value = float(loss_host[0])

    ar_info = None
    if world > 1:
        # the collective by itself (ranks in lock-step, CUDA events, max over ranks): what it adds to a step
        def alone(fn, n=50):
            for _ in range(5):
                fn()
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(n):
                fn()
            a1.record()
            torch.cuda.synchronize()
            return 1e3 * a0.elapsed_time(a1) / n
        gw, gb = st["weight"].grad.clone(), st["bias"].grad.clone()
        us_nccl = alone(lambda: dist.all_reduce(bucket))
        us_peer = alone(lambda: peer_ar.all_reduce_mean_([gw, gb])) if peer_ar is not None else None
        t = torch.tensor([us_nccl, us_peer or 0.0], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        comp = [torch.zeros(1, device=dev, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(comp, torch.tensor([compute_only_ms], device=dev, dtype=torch.float64))
        q = lambda v, f: v[min(len(v) - 1, int(f * len(v)))]
        ar_info = {"impl": "peer-memory one-shot kernel inside the step's CUDA graph" if peer_ar is not None
                   else "NCCL all_reduce of one flat bucket, issued from the host after the graph",
                   "floats": bucket.numel(), "us_alone_nccl": float(t[0]),
                   "us_alone_peer": float(t[1]) if peer_ar is not None else None, "note": ar_note,
                   "breakdown": {
                       "what": "rank-0 per-step distribution of the timed steps (ms); every rank's step WITHOUT the "
                               "collective (own graph, nobody to wait for); the collective alone is us_alone_*: "
                               "step - (compute + collective) = waiting for the slowest rank",
                       "step_ms_p10_p50_p90_max": [q(step_times, 0.1), q(step_times, 0.5), q(step_times, 0.9), step_times[-1]],
                       "compute_only_ms_per_rank": [float(c[0]) for c in comp],
                       "compute_only_ms_p50_p90_max_rank0": [q(step_times_local, 0.5), q(step_times_local, 0.9), step_times_local[-1]]}}
        t = torch.tensor([total_ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms = float(t[0]), float(t[1])
        cells_t = torch.tensor([cells], device=dev, dtype=torch.float64)
        dist.all_reduce(cells_t)
        total_cells = float(cells_t[0])
    else:
        total_cells = float(cells)

    # ---- per-kernel timing through the C ABI (rank 0): roofline -------------------------------
    kernels, roofline = {}, None
    if rank == 0:
        kernels = per_kernel(lib, st, mode, GEMMS[gemm], det, B, T, U1, V, H, cells, flush_buf)
        peak, peak_src = hbm_peak()
        for k in kernels.values():
            k["GBps"] = k["bytes"] / (k["us"] * 1e-6) / 1e9
            k["frac_hbm"] = k["GBps"] / peak
        # the roofline object is about the kernel north_star's target names: the alpha/beta lattice sweep
        top = kernels["lattice_sweep_kernel"]
        slowest = max((k for k in kernels.values() if k.get("ours")), key=lambda k: k["us"])
        roofline = {"kernel": top["name"], "bound": "hbm", "achieved": top["GBps"], "peak": peak,
                    "unit": "GB/s", "frac": top["frac_hbm"], "traffic": ncu_traffic(top["name"], c, args),
                    "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": top["bytes"], "us_per_launch": top["us"],
                    "slowest_kernel_of_step": {"kernel": slowest["name"], "us": slowest["us"], "frac": slowest["frac_hbm"]},
                    "note": "at the workload batch the sweep is 2B independent chains of T+U dependent steps "
                            "(latency-bound, DESIGN.md section 6); saturating_batch is the same kernel with enough "
                            "utterances to fill the GPU"}
        sat = sweep_saturating(lib, st, B, T, U1, flush_buf)
        sat["GBps"] = sat["bytes"] / (sat["us"] * 1e-6) / 1e9
        sat["frac"] = sat["GBps"] / peak
        roofline["saturating_batch"] = sat

    used_peer = peer_ar is not None
    if used_peer:
        peer_ar.close()  # collective: every rank is here
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # our kernels per step: concat_gelu = weight split + projection + factor rows + lse + sweep + grad
    # + projection backward (+ slab reduction); add_tanh = weight convert + lse + sweep + (weight
    # convert +) grad
    launches = {"concat_gelu": 7 + int(det), "add_tanh": 5 if gemm == "bf16" else 3}[mode] + int(used_peer)
    line = {
        "metric": METRIC, "value": total_cells * args.steps / (total_ms * 1e-3), "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": ("f32" if gemm == "fp32" else "bf16 GEMM / f32 lattice") +
                                      ("" if act_dtype == torch.float32 else f" ({args.act_dtype} activations, fp32 arithmetic)"),
        "data": "synthetic", "config": config,
        "utterances_per_s": B * world * args.steps / (total_ms * 1e-3),
        "cells_per_step": total_cells, "loss": loss_value,
        "cuda_graph": graph is not None,
        "step_ms_p10_p50_p90_max": [step_times[int(0.1 * len(step_times))], step_times[len(step_times) // 2],
                                    step_times[min(len(step_times) - 1, int(0.9 * len(step_times)))], step_times[-1]],
        "clocks": clocks.summary(),
        "e2e": {"value": total_cells * args.steps / (e2e_ms * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / args.steps,
                "wall_ms_per_step": 1e3 * e2e_wall / args.steps,
                "pipeline": "one pinned slab -> ONE H2D copy per step; upload of step i+1 overlaps step i (copy stream, double-buffered); loss read every step"},
        "gpu_launches": launches * args.steps,
        "roofline": roofline, "kernels": kernels,
    }
    if ar_info is not None:
        line["allreduce"] = ar_info
        line["config"]["parallelism"] = f"dp{world} (utterances sharded, fc grads all-reduced: {ar_info['impl']})"
    if world == 1 and not args.no_gpu_baseline:
        torch.cuda.empty_cache()
        line["gpu_baseline"] = gpu_baseline(st, mode, cells)
        if "value" in line["gpu_baseline"]:
            line["gpu_baseline"]["ours_over_baseline"] = line["value"] / line["gpu_baseline"]["value"]
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = time_cpu(host32, mode, args.cpu_seconds)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# BASELINE cfg 5: the full training step under DDP
CFG5_ENCODER = dict(input_size=80, hidden_size=1024, output_size=512, num_layers=8, rnn_type="lstm", dropout=0.0,
                    bidirectional=True)   # reference config/config.json:3-11 with LSTM cells (BASELINE configs[4])
CFG5_PREDICTOR = dict(embedding_size=73, hidden_size=1024, output_size=512, num_layers=2, rnn_type="lstm", dropout=0.0)


def run_cfg5(args):
    """Full RNNTransducer training step (reference model.py:52-60 + 110-126 under train.py:45-48): LSTM
    encoder / predictor (cuDNN, library) -> fused joint + RNN-T loss (ours) -> backward -> DDP bucketed
    all-reduce overlapping the cuDNN backward -> AdamW + OneCycleLR.  Per rank the cfg-2 batch (B=32, T=400
    log-mel frames of 80 bins, U=80, V=73, joint width 512); metric = lattice cells/s of the whole step."""
    import torch
    import torch.distributed as dist

    import rnntransducer_b200 as rb
    from rnntransducer_b200 import _lib, synthetic
    from rnntransducer_b200.training import synthetic_training_batch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl ours) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    c, gemm, config = workload_config(args, world)
    B, T, U, V, H = c["B"], c["T"], c["U"], c["V"], c["H"]
    torch.manual_seed(1234)  # same initial parameters on every rank (DDP broadcasts rank 0's anyway)
    step_mod = rb.RNNTransducerStep(dict(CFG5_PREDICTOR), dict(CFG5_ENCODER), dict(num_classes=V), blank_token_id=0,
                                    mode=args.mode, gemm=gemm, deterministic=bool(args.deterministic)).to(dev).train()
    n_params = sum(p.numel() for p in step_mod.parameters())
    model = step_mod
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(step_mod, device_ids=[local_rank], gradient_as_bucket_view=True)
    total_steps = max(args.warmup, 3) + 2 * args.steps + 8
    opt, sched = rb.configure_optimizers(step_mod, learning_rate=1e-4, weight_decay=1e-2, total_steps=total_steps)
    host = synthetic_training_batch(B, T, U, CFG5_ENCODER["input_size"], V, ragged=args.ragged, seed=1239 + rank)
    cells = synthetic.count_cells(host[2], host[6])
    tensors = [i for i, x in enumerate(host) if torch.is_tensor(x)]
    pinned = {i: host[i].pin_memory() for i in tensors}
    dev_batch = list(host)
    for i in tensors:
        dev_batch[i] = host[i].to(dev)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {}

    def train_step(batch):
        opt.zero_grad(set_to_none=True)
        loss = model(*batch)
        loss.backward()          # DDP all-reduces its buckets while cuDNN is still in the encoder's backward
        opt.step()
        sched.step()
        out["loss"] = loss.detach()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local_rank).start()  # nvmlInit and the sampling thread before the barrier (see run_ours)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for _ in range(max(args.warmup, 3)):
        train_step(dev_batch)
    barrier()
    with clocks:
        for e0, e1 in ev:
            flush_buf.zero_()
            e0.record()
            train_step(dev_batch)
            e1.record()
        torch.cuda.synchronize()
    total_ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    barrier()

    # end to end: the collate's batch arrives in pinned host memory every step, the loss is read back
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()
    h2d = sum(pinned[i].numel() * pinned[i].element_size() for i in tensors)

    def e2e_step():
        batch = list(host)
        for i in tensors:
            batch[i] = pinned[i].to(dev, non_blocking=True)
        train_step(batch)
        loss_host.copy_(out["loss"].reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = e0.elapsed_time(e1)
    barrier()
    total_cells = float(cells)
    if world > 1:
        t = torch.tensor([total_ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms = float(t[0]), float(t[1])
        ct = torch.tensor([cells], device=dev, dtype=torch.float64)
        dist.all_reduce(ct)
        total_cells = float(ct[0])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # roofline of OUR part of the step: the joint + loss kernels on encoder / predictor outputs of this shape
    from rnntransducer_b200.joint_add_tanh import GEMMS
    syn = {k: v.to(dev) for k, v in synthetic.make_batch(B, T, U, V, H, mode=args.mode, ragged=args.ragged, seed=1239).items()}
    kernels = per_kernel(lib, syn, args.mode, GEMMS[gemm], bool(args.deterministic), B, T, U + 1, V, H, cells, flush_buf)
    peak, peak_src = hbm_peak()
    for k in kernels.values():
        k["GBps"] = k["bytes"] / (k["us"] * 1e-6) / 1e9
        k["frac_hbm"] = k["GBps"] / peak
    top = kernels["lattice_sweep_kernel"]
    ours_us = sum(k["us"] for k in kernels.values() if k.get("ours"))
    config.update({"workload": f"BASELINE cfg5: full training step, per GPU B={B} T={T} (80-bin log-mel) U={U} V={V}; "
                               f"LSTM encoder 8x1024 bidirectional -> 512, LSTM predictor 2x1024 -> 512, fused joint + RNN-T "
                               f"loss, AdamW + OneCycleLR, torch DDP ({n_params / 1e6:.0f} M parameters)",
                   "parallelism": f"ddp{world} (torch DistributedDataParallel, NCCL bucketed all-reduce overlapping backward)",
                   "parameters": n_params, "precision": "fp32 parameters, cuDNN RNN with TF32 tensor cores (torch default)"})
    line = {
        "metric": METRIC, "value": total_cells * args.steps / (total_ms * 1e-3), "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (TF32 cuDNN RNN)", "data": "synthetic", "config": config,
        "utterances_per_s": B * world * args.steps / (total_ms * 1e-3), "cells_per_step": total_cells,
        "loss": float(out["loss"]), "cuda_graph": False, "clocks": clocks.summary(),
        "e2e": {"value": total_cells * args.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": (7 + int(bool(args.deterministic))) * args.steps,
        "roofline": {"kernel": top["name"], "bound": "hbm", "achieved": top["GBps"], "peak": peak, "unit": "GB/s",
                     "frac": top["frac_hbm"], "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": top["bytes"], "us_per_launch": top["us"]},
        "kernels": kernels,
        "joint_loss_share_of_step": ours_us * 1e-3 / (total_ms / args.steps),
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def reference_joint_eager(enc, dec, weight, bias, mode):
    """The reference's JointNet.joint as it runs on its GPU (networks/transducer.py:54-71: unsqueeze ->
    repeat x2 -> cat -> GELU(tanh) -> Linear), materialising every [B,T,U1,2H] intermediate; add_tanh:
    torchaudio's _Joiner expression.  Baseline only -- the product path never calls this."""
    import torch
    import torch.nn.functional as F
    if mode == "concat_gelu":
        T, U1 = enc.size(1), dec.size(1)
        e = enc.unsqueeze(2).repeat([1, 1, U1, 1])
        d = dec.unsqueeze(1).repeat([1, T, 1, 1])
        out = F.gelu(torch.cat((e, d), dim=-1), approximate="tanh")
    else:
        out = torch.tanh(enc.unsqueeze(2) + dec.unsqueeze(1))
    return F.linear(out, weight, bias)


def gpu_baseline(st, mode, cells, iters=10, warmup=3):
    """The existing-GPU-kernel bar: the reference's eager joint feeding
    torchaudio.functional.rnnt_loss on CUDA (its shipped loss, model.py:28-31,57; SIMT kernels compiled
    for sm_100 in libtorchaudio), fwd+bwd on the SAME device-resident batch, CUDA events.  fp32 like our
    arm, and once more under fp16 autocast (scripts/run_train.sh:32 --precision=16)."""
    import torch
    try:
        import torchaudio
    except Exception as e:
        return {"unavailable": f"torchaudio import failed: {e!r}"}
    leaves = {k: st[k].detach().float().clone().requires_grad_(True) for k in ("enc", "dec", "weight", "bias")}
    lab, al, ll = st["labels"].contiguous(), st["act_lens"].contiguous(), st["label_lens"].contiguous()

    def step(fp16):
        for v in leaves.values():
            v.grad = None
        with torch.autocast("cuda", dtype=torch.float16, enabled=fp16):
            logits = reference_joint_eager(leaves["enc"], leaves["dec"], leaves["weight"], leaves["bias"], mode)
        loss = torchaudio.functional.rnnt_loss(logits, lab, al, ll, blank=0, reduction="mean")
        loss.backward()
        return loss

    out = {"what": "eager reference JointNet.joint (repeat/cat/GELU/Linear) + torchaudio.functional.rnnt_loss "
                   "on CUDA, fwd+bwd, same batch resident in HBM, CUDA events", "unit": UNIT}
    for name, fp16 in (("fp32", False), ("fp16_autocast", True)):
        try:
            torch.cuda.reset_peak_memory_stats()
            for _ in range(warmup):
                loss = step(fp16)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                loss = step(fp16)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            r = {"value": cells / (ms * 1e-3), "ms_per_step": ms, "loss": float(loss),
                 "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9, "iters": iters}
        except Exception as e:  # out of memory on the eager intermediates, missing CUDA op, ...
            r = {"unavailable": repr(e)[:300]}
        if name == "fp32":
            out.update(r)
        else:
            out[name] = r
        for v in leaves.values():
            v.grad = None
        torch.cuda.empty_cache()
    # the loss alone, like for like on the SAME dense logits (row A8: RNNTLoss as a drop-in for torchaudio's):
    # torchaudio.functional.rnnt_loss vs rnntransducer_b200.rnnt_loss, fwd+bwd w.r.t. the logits
    try:
        import rnntransducer_b200 as rb
        with torch.no_grad():
            logits = reference_joint_eager(leaves["enc"], leaves["dec"], leaves["weight"], leaves["bias"], mode)
        logits.requires_grad_(True)

        def time_loss(fn):
            for _ in range(warmup):
                logits.grad = None
                fn().backward()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                logits.grad = None
                loss = fn()
                loss.backward()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters, float(loss)
        ta_ms, ta_loss = time_loss(lambda: torchaudio.functional.rnnt_loss(logits, lab, al, ll, blank=0, reduction="mean"))
        our_ms, our_loss = time_loss(lambda: rb.rnnt_loss(logits, lab, al, ll, 0, "mean", warp_compat=False))
        out["loss_only_dense_logits"] = {"torchaudio_cuda_ms": ta_ms, "ours_dense_ms": our_ms, "speedup": ta_ms / our_ms,
                                         "loss_torchaudio": ta_loss, "loss_ours": our_loss,
                                         "logits_GB": logits.numel() * 4 / 1e9}
        del logits
    except Exception as e:
        out["loss_only_dense_logits"] = {"unavailable": repr(e)[:300]}
    torch.cuda.empty_cache()
    return out


def sweep_saturating(lib, st, B, T, U1, flush_buf, iters=20):
    """The lattice sweep alone at a batch that fills the GPU (the workload's shape and lengths, Bs ~ 2048
    utterances, <= 2 GiB of lp2), timed like per_kernel (L2 flushed, CUDA events)."""
    import torch
    from rnntransducer_b200 import _lib
    dev = st["enc"].device
    # ~2048 utterances (about nine waves of CTAs: a 2.3-wave run at B = 512 loses a fifth to its last, third-full
    # wave), bounded by 2 GiB of lp2
    rep = max(1, min(2048 // max(B, 1), (2 << 30) // max(B * T * U1 * 8, 1)))
    Bs = B * rep
    f32 = dict(device=dev, dtype=torch.float32)
    stream = torch.cuda.current_stream().cuda_stream
    p = lambda t: t.data_ptr()
    al, ll = st["act_lens"].repeat(rep).contiguous(), st["label_lens"].repeat(rep).contiguous()
    # realistic factors: per-cell log-probs of a V-way softmax over N(0,1) logits
    torch.manual_seed(0)
    lp2 = (torch.randn(Bs, T, U1, 2, **f32) - 4.8).clamp_(max=-0.05).contiguous()
    alpha, beta = (torch.empty(Bs, T, U1, device=dev, dtype=torch.int32) for _ in range(2))
    costs = torch.empty(Bs, **f32)
    fn = lambda: _lib.check(lib.rnntb200_lattice_sweep(p(lp2), p(al), p(ll), Bs, T, U1, p(alpha), p(beta),
                                                       p(costs), None, stream))
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush_buf.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    cells = int((al.long() * (ll.long() + 1)).sum().item())
    return {"B": Bs, "us": 1e3 * tot / iters, "bytes": 24 * cells, "cells": cells}


def per_kernel(lib, st, mode, gemm, det, B, T, U1, V, H, cells, flush_buf, iters=20):
    """Times each kernel of the step by itself (CUDA events on the launching stream, L2 flushed
    between launches).  `bytes` = algorithmic bytes per launch (DESIGN.md / SURVEY 8(d))."""
    import torch
    import torch.nn.functional as F
    from rnntransducer_b200 import _lib
    dev = st["enc"].device
    f32 = dict(device=dev, dtype=torch.float32)
    stream = torch.cuda.current_stream().cuda_stream
    p = lambda t: t.data_ptr()
    lp2 = torch.empty(B, T, U1, 2, **f32)
    lse = torch.empty(B, T, U1, **f32)
    alpha, beta = (torch.empty(B, T, U1, device=dev, dtype=torch.int32) for _ in range(2))  # e16m16 planes
    costs, gcosts = torch.empty(B, **f32), torch.full((B,), 1.0 / B, **f32)
    lab, al, ll = st["labels"], st["act_lens"], st["label_lens"]
    res = {}

    def bench(name, fn, nbytes, ours=True):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(iters):
            flush_buf.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            tot += e0.elapsed_time(e1)
        res[name] = {"name": name, "us": 1e3 * tot / iters, "bytes": int(nbytes), "ours": ours}

    with torch.no_grad():
        enc, dec, w, b = st["enc"].detach(), st["dec"].detach(), st["weight"].detach(), st["bias"].detach()
        if mode == "concat_gelu":
            He = enc.shape[-1]
            xdt = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}[enc.dtype]
            xb = enc.element_size()
            proj = lambda: (F.linear(F.gelu(enc.float(), approximate="tanh"), w[:, :He], b),
                            F.linear(F.gelu(dec.float(), approximate="tanh"), w[:, He:]))
            penc, pdec = proj()
            penc, pdec = penc.contiguous(), pdec.contiguous()
            d_penc, d_pdec = torch.empty_like(penc), torch.empty_like(pdec)
            ws_bytes = lib.rnntb200_joint_cg_bwd_workspace_bytes(B, T, U1, V, int(det))
            ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
            io = 4 * V * B * (T + U1)
            pws_bytes = lib.rnntb200_joint_cg_project_workspace_bytes(V, He, dec.shape[-1])
            if pws_bytes:
                pws = torch.empty(pws_bytes, dtype=torch.uint8, device=dev)
                bench("proj_tc_kernel", lambda: _lib.check(lib.rnntb200_joint_cg_project(
                    p(enc), p(dec), xdt, p(w), p(b), B * T, B * U1, He, dec.shape[-1], V, p(penc), p(pdec), p(pws),
                    pws_bytes, stream)), xb * (enc.numel() + dec.numel()) + 4 * w.numel() + io)
                res["proj_tc_kernel"]["flops"] = 2.0 * V * (He * B * T + dec.shape[-1] * B * U1)
            else:
                bench("torch_projections(gelu+linear x2, library)", proj,
                      4 * (enc.numel() + dec.numel() + w.numel()) + io, ours=False)
            fac_bytes = lib.rnntb200_joint_cg_factors_bytes(B, T, U1, V)
            fac = torch.empty(max(fac_bytes, 16), dtype=torch.uint8, device=dev)
            # factor-rows kernel + cell kernel (two launches, timed together)
            bench("cg_lse_kernel", lambda: _lib.check(lib.rnntb200_joint_cg_logprobs(
                p(penc), p(pdec), p(lab), p(al), p(ll), B, T, U1, V, 0, p(lp2), p(lse),
                p(fac) if fac_bytes else None, fac_bytes, stream)), 12 * cells + io)
            bench("lattice_sweep_kernel", lambda: _lib.check(lib.rnntb200_lattice_sweep(
                p(lp2), p(al), p(ll), B, T, U1, p(alpha), p(beta), p(costs), None, stream)), 24 * cells)
            bench("cg_grad_kernel", lambda: _lib.check(lib.rnntb200_joint_cg_bwd(
                p(penc), p(pdec), p(lab), p(al), p(ll), B, T, U1, V, 0, p(lse), p(alpha), p(beta),
                p(gcosts), p(d_penc), p(d_pdec), int(det), p(ws), ws_bytes,
                p(fac) if fac_bytes else None, fac_bytes, stream)), 12 * cells + 2 * io)
            bws_bytes = lib.rnntb200_joint_cg_project_bwd_workspace_bytes(V, He, dec.shape[-1])
            if bws_bytes:
                bws = torch.empty(bws_bytes, dtype=torch.uint8, device=dev)
                d_enc, d_dec, d_w, d_b = (torch.empty_like(t) for t in (enc, dec, w, b))
                bench("proj_tc_bwd_kernel", lambda: _lib.check(lib.rnntb200_joint_cg_project_bwd(
                    p(enc), p(dec), xdt, p(w), p(d_penc), p(d_pdec), B * T, B * U1, He, dec.shape[-1], V, p(d_enc),
                    p(d_dec), p(d_w), p(d_b), p(bws), bws_bytes, 0, stream)),
                    2 * xb * (enc.numel() + dec.numel()) + 8 * w.numel() + io)
                res["proj_tc_bwd_kernel"]["flops"] = 4.0 * V * (He * B * T + dec.shape[-1] * B * U1)
        else:
            d_enc, d_dec = torch.empty_like(enc), torch.empty_like(dec)
            d_w, d_b = torch.empty_like(w), torch.empty_like(b)
            io = 4 * H * B * (T + U1) + 4 * V * (H + 1)
            ws_bytes = lib.rnntb200_joint_at_workspace_bytes(V, H, gemm)
            ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
            bench("at_lse_kernel", lambda: _lib.check(lib.rnntb200_joint_at_logprobs(
                p(enc), p(dec), p(w), p(b), gemm, p(lab), p(al), p(ll), B, T, U1, V, H, 0, p(lp2),
                p(lse), p(ws), ws_bytes, stream)), 12 * cells + io)
            bench("lattice_sweep_kernel", lambda: _lib.check(lib.rnntb200_lattice_sweep(
                p(lp2), p(al), p(ll), B, T, U1, p(alpha), p(beta), p(costs), None, stream)), 24 * cells)
            bench("at_grad_kernel", lambda: _lib.check(lib.rnntb200_joint_at_bwd(
                p(enc), p(dec), p(w), p(b), gemm, p(lab), p(al), p(ll), B, T, U1, V, H, 0, p(lp2), p(lse),
                p(alpha), p(beta), p(gcosts), p(d_enc), p(d_dec), p(d_w), p(d_b), p(ws), ws_bytes, stream)),
                12 * cells + 2 * io)
            for k in ("at_lse_kernel", "at_grad_kernel"):
                passes = 1 if k == "at_lse_kernel" else 3  # fwd | recompute + dgrad + wgrad
                res[k]["flops"] = 2.0 * H * V * cells * passes
                res[k]["TFLOPs"] = res[k]["flops"] / (res[k]["us"] * 1e-6) / 1e12
    return res


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the real stdout; everything else any library prints to fd 1 during
    the run (e.g. NCCL's version banner) was redirected to stderr in main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # stray prints of libraries -> stderr
    if args.impl == "reference":
        run_reference(args)
    elif args.cfg == 5:
        run_cfg5(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
