"""CPU, gloo, world_size 2: the N>1 host logic (utterance sharding + gradient all-reduce) gives the
single-process result.  No GPU here, so each rank evaluates its shard with the CPU oracle; the GPU
kernels are per-utterance independent, which tests/test_gpu_* establish (sub-batch == full batch)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rnntransducer_b200 import distributed as rd
from rnntransducer_b200 import synthetic


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _step(batch):
    from oracle import joint_ref
    r = joint_ref.joint_loss_fwd_bwd(batch["enc"], batch["dec"], batch["weight"], batch["bias"],
                                     batch["labels"].numpy(), batch["act_lens"].numpy(),
                                     batch["label_lens"].numpy(), 0, "mean", "concat_gelu", num_threads=1)
    return r


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    full = synthetic.make_batch(4, 12, 5, 9, 8, ragged=True, seed=99)
    shard = rd.shard_batch(full, rank, world)
    assert shard["enc"].shape[0] == 2 and shard["weight"] is full["weight"]
    r = _step(shard)
    grads = [torch.from_numpy(r["d_weight"].copy()), torch.from_numpy(r["d_bias"].copy())]
    loss = torch.from_numpy(np.asarray(r["loss"], dtype=np.float32).copy())
    rd.allreduce_mean_(grads + [loss])
    # per-utterance gradients stay local: gather them back in shard order to compare
    d_enc = [torch.zeros(2, 12, 8) for _ in range(world)]
    dist.all_gather(d_enc, torch.from_numpy(r["d_enc"].copy()))
    if rank == 0:
        out.put((grads[0].numpy(), grads[1].numpy(), float(loss[0]), [t.numpy() for t in d_enc]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_step_matches_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    d_w, d_b, loss, d_enc = out.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = synthetic.make_batch(4, 12, 5, 9, 8, ragged=True, seed=99)
    ref = _step(full)
    np.testing.assert_allclose(loss, float(ref["loss"][0]), rtol=1e-6)
    np.testing.assert_allclose(d_w, ref["d_weight"], atol=1e-6)
    np.testing.assert_allclose(d_b, ref["d_bias"], atol=1e-6)
    # rank r holds utterances r, r+2: its d_enc (of the rank-local mean over 2) is 2x the global one
    for r in range(world):
        idx = rd.shard_indices(4, r, world).numpy()
        np.testing.assert_allclose(d_enc[r] / world, ref["d_enc"][idx], atol=1e-6)


def test_shard_indices():
    assert rd.shard_indices(7, 1, 3).tolist() == [1, 4]
    assert rd.shard_indices(8, 3, 4).tolist() == [3, 7]
