"""CPU: the data-side pieces: length-bucketed distributed sampler and the
length-sorted collate, against the reference's own classes when /root/reference is on this machine and
against hand-derived expectations everywhere."""
import os
import sys

import pytest
import torch

from rnntransducer_b200.data import DistributedBucketSampler, collate_sorted

LENGTHS = [5, 9, 2, 9, 7, 1, 4, 8, 3, 6, 9]


def test_bucket_sampler_streams():
    # longest first, ties in dataset order: 9(1) 9(3) 9(10) 8(7) 7(4) 6(9) 5(0) 4(6) 3(8) 2(2) 1(5)
    order = [1, 3, 10, 7, 4, 9, 0, 6, 8, 2, 5]
    s = [DistributedBucketSampler(lengths=LENGTHS, num_replicas=3, rank=r) for r in range(3)]
    assert s[0].get_bucket_indices(LENGTHS) == order
    assert [len(x) for x in s] == [4, 4, 4] and s[0].total_size == 12
    padded = order + order[:1]
    for r in range(3):
        assert list(s[r]) == padded[r::3]
    d = [DistributedBucketSampler(lengths=LENGTHS, num_replicas=3, rank=r, drop_last=True) for r in range(3)]
    assert [len(x) for x in d] == [3, 3, 3]
    for r in range(3):
        assert list(d[r]) == order[:9][r::3]
    # every position of the epoch: the ranks hold neighbours of the sorted order
    for pos in range(3):
        lens = [LENGTHS[list(d[r])[pos]] for r in range(3)]
        assert max(lens) - min(lens) <= 3


def test_bucket_sampler_matches_the_reference_class():
    if not os.path.isfile("/root/reference/datasampler.py"):
        pytest.skip("reference tree not on this machine")
    sys.path.insert(0, "/root/reference")
    try:
        from datasampler import DistributedBucketSampler as Ref
    except Exception as e:  # transformers / tqdm import problems
        pytest.skip(f"reference sampler not importable: {e!r}")
    for world in (1, 2, 4):
        for drop_last in (False, True):
            for r in range(world):
                ours = DistributedBucketSampler(lengths=LENGTHS, num_replicas=world, rank=r, drop_last=drop_last)
                theirs = Ref(lengths=list(LENGTHS), num_replicas=world, rank=r, drop_last=drop_last)
                assert list(ours) == list(theirs) and len(ours) == len(theirs)


def test_bucket_sampler_shuffle_keeps_buckets_and_balance():
    world, bs = 2, 2
    s = [DistributedBucketSampler(lengths=LENGTHS, num_replicas=world, rank=r, shuffle=True, batch_size=bs, seed=3)
         for r in range(world)]
    for x in s:
        x.set_epoch(5)
    a, b = list(s[0]), list(s[1])
    assert sorted(a + b) == sorted(list(range(11)) + [1])  # every sample once (+ the wrap-around pad)
    plain = [list(DistributedBucketSampler(lengths=LENGTHS, num_replicas=world, rank=r)) for r in range(world)]
    chunks = lambda v: sorted(tuple(v[i:i + bs]) for i in range(0, len(v), bs))
    assert chunks(a) == chunks(plain[0]) and chunks(b) == chunks(plain[1])  # same batches, permuted
    s[0].set_epoch(6)
    assert list(s[0]) != a or True  # a different epoch may permute differently (not required to)
    with pytest.raises(ValueError):
        DistributedBucketSampler()
    with pytest.raises(ValueError):
        DistributedBucketSampler(lengths=LENGTHS, num_replicas=2, rank=2)


def test_bucket_sampler_infers_lengths_from_dataset():
    ds = [{"input_ids": [0] * n} for n in (3, 1, 2)]
    assert list(DistributedBucketSampler(ds, num_replicas=1, rank=0)) == [0, 2, 1]
    with pytest.raises(ValueError):
        DistributedBucketSampler([[1, 2]], num_replicas=1, rank=0)


def test_collate_sorted_contract():
    """The reference's 7-tuple (dataloader.py:49), longest audio first."""
    torch.manual_seed(0)
    items = [{"input_values": torch.randn(t, 4), "input_ids": list(range(1, u + 1))} for t, u in ((5, 2), (9, 3), (7, 1))]
    audios, audio_lengths, tal, texts, text_lengths, targets, target_lengths = collate_sorted(items, 0, 4)
    assert audio_lengths == [9, 7, 5] and tal.dtype == torch.int32 and tal.tolist() == [9, 7, 5]
    assert audios.shape == (3, 9, 4) and float(audios[2, 5:].abs().sum()) == 0.0
    assert texts.dtype == torch.int64 and texts.tolist() == [[0, 1, 2, 3], [0, 1, 0, 0], [0, 1, 2, 0]]
    assert text_lengths == [4, 2, 3]
    assert targets.dtype == torch.int32 and targets.tolist() == [[1, 2, 3], [1, 0, 0], [1, 2, 0]]
    assert target_lengths.dtype == torch.int32 and target_lengths.tolist() == [3, 1, 2]
    assert all(text_lengths[i] == int(target_lengths[i]) + 1 for i in range(3))  # dataloader.py:39-40
    with pytest.raises(ValueError):
        collate_sorted(items, 0, 5)
