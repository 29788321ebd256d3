"""GPU: the peer-memory gradient all-reduce (csrc/comm.cu) against torch.distributed's all_reduce.

World 1 (one process) checks the kernel's copy / sum / scale plumbing on any box; world 2 needs two GPUs
(two ranks spinning on each other cannot share one device) and is skipped elsewhere -- it runs in the
round's --gpus 2 visit (scripts/gpu_multi2.sh)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl" if world > 1 else "gloo", rank=rank, world_size=world)
    from rnntransducer_b200.comm import PeerAllReduce
    ar = PeerAllReduce(74752 + 73)
    ok = True
    g = torch.Generator().manual_seed(100 + rank)
    for step in range(7):  # odd count: both parity halves, reuse of each
        w = torch.randn(73, 1024, generator=g).cuda()
        b = torch.randn(73, generator=g).cuda()
        want_w, want_b = w.clone(), b.clone()
        if world > 1:
            dist.all_reduce(want_w)
            dist.all_reduce(want_b)
        ar.all_reduce_mean_([w, b])
        torch.cuda.synchronize()
        ok &= torch.allclose(w, want_w / world, atol=1e-6) and torch.allclose(b, want_b / world, atol=1e-6)
    # inside a CUDA graph: parameter-free launch, replayed
    w = torch.full((73, 1024), float(rank + 1), device="cuda")
    b = torch.full((73,), 2.0 * (rank + 1), device="cuda")
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        ar.all_reduce_mean_([w, b], average=False)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    w.fill_(float(rank + 1))
    with torch.cuda.graph(graph):
        ar.all_reduce_mean_([w, b], average=False)
    total = world * (world + 1) / 2
    for _ in range(3):
        w.fill_(float(rank + 1))
        graph.replay()
        torch.cuda.synchronize()
        ok &= bool((w == total).all())
    out.put((rank, ok))
    ar.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2])
def test_peer_allreduce_matches_torch_distributed(cuda_lib, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(out.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(res.values()), res
