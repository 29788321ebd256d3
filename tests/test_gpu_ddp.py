"""GPU, world_size 2: the full training step (reference model.py:52-60 under train.py:45-48's DDP) -- the
LSTM encoder / predictor + fused joint + loss wrapped in torch DistributedDataParallel -- gives, after the
gradient all-reduce, the gradients of ONE process on the concatenated batch.

Two ranks over NCCL, one GPU each: skipped on a one-GPU box (it runs in the round's --gpus 2 visit,
scripts/gpu_multi2.sh; the host-side sharding logic is covered on CPU by tests/test_ddp_cpu.py).  The
single-process half of the statement -- the step module, its optimizer, loss going down -- runs anywhere."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

EP = dict(input_size=16, hidden_size=24, output_size=128, num_layers=2, rnn_type="lstm", dropout=0.0, bidirectional=True)
DP = dict(embedding_size=29, hidden_size=24, output_size=128, num_layers=1, rnn_type="lstm", dropout=0.0)
B, T, U = 6, 40, 9


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make(seed=11):
    import rnntransducer_b200 as rb
    from rnntransducer_b200.training import synthetic_training_batch
    torch.manual_seed(seed)
    step = rb.RNNTransducerStep(dict(DP), dict(EP), dict(num_classes=29), blank_token_id=0, deterministic=True)
    batch = synthetic_training_batch(B, T, U, 16, 29, ragged=True, seed=seed)
    return step, batch


def _shard(batch, idx):
    """Rank-local batch as the collate would have built it: padded to ITS longest audio / text."""
    audios, al, tal, texts, tl, targets, tgl = batch
    sel = torch.tensor(idx)
    a_max, t_max = max(al[i] for i in idx), max(tl[i] for i in idx)
    return (audios[sel][:, :a_max].contiguous(), [al[i] for i in idx], tal[sel], texts[sel][:, :t_max].contiguous(),
            [tl[i] for i in idx], targets[sel][:, :t_max - 1].contiguous(), tgl[sel])


def _to_dev(batch, dev):
    return tuple(x.to(dev) if torch.is_tensor(x) else x for x in batch)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    step, batch = _make()
    step = step.to(dev).train()
    ddp = torch.nn.parallel.DistributedDataParallel(step, device_ids=[dev.index])
    loss = ddp(*_to_dev(_shard(batch, list(range(rank, B, world))), dev))
    loss.backward()
    torch.cuda.synchronize()
    if rank == 0:
        out.put((float(loss), {n: p.grad.detach().cpu() for n, p in step.named_parameters()}))
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_step_matches_single_process_on_the_concatenated_batch(cuda_lib):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (NCCL refuses two ranks on one device)")
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    loss0, grads = out.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    step, batch = _make()
    step = step.cuda().train()
    loss = step(*_to_dev(batch, torch.device("cuda")))
    loss.backward()
    # rank 0's loss is the mean over ITS shard; the averaged gradients are those of the global mean
    for n, p in step.named_parameters():
        torch.testing.assert_close(grads[n], p.grad.cpu(), atol=2e-5, rtol=1e-4, msg=n)
    assert loss.shape == (1,) and abs(loss0 - float(loss)) < 0.5 * abs(float(loss))


def test_training_step_module_and_optimizer(cuda_lib):
    """RNNTransducerStep = model.py:52-57; configure_optimizers = model.py:110-126.  A few optimizer steps
    on one batch must reduce the loss, the lazy handle must be what the loss consumed (no dense logits)."""
    import rnntransducer_b200 as rb
    step, batch = _make(seed=5)
    step = step.cuda().train()
    assert {k.split(".")[0] for k in step.state_dict()} == {"jointnet"}
    opt, sched = rb.configure_optimizers(step, learning_rate=3e-3, weight_decay=1e-2, total_steps=12, warmup_ratio=0.25,
                                         final_div_factor=10.0)
    dev_batch = _to_dev(batch, torch.device("cuda"))
    losses = []
    for _ in range(12):
        opt.zero_grad(set_to_none=True)
        loss = step(*dev_batch)
        loss.backward()
        opt.step()
        sched.step()
        losses.append(float(loss))
    assert losses[-1] < 0.8 * losses[0], losses
    assert isinstance(step.jointnet(dev_batch[0], dev_batch[1], dev_batch[3], dev_batch[4]), rb.JointLogits)
