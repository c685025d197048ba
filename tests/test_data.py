"""Unit tier (CPU): Partition / DataPartitioner / partition_dataset."""
import pytest
import torch
from hypothesis import given, settings, strategies as st

import dist_tuto.pth_b200 as b2
from dist_tuto.pth_b200 import data as D


def test_golden_partition():
    # probed on the reference with py3.12
    assert D.DataPartitioner(list(range(10))).partitions == [[2, 8, 3, 5, 6, 4, 9], [0, 1], [7]]


def test_partition_view():
    base = [x * x for x in range(20)]
    p = D.Partition(base, [3, 1, 7])
    assert len(p) == 3 and [p[i] for i in range(3)] == [9, 1, 49]


@pytest.mark.parametrize("world", [1, 2, 3, 4, 6, 7, 8])
def test_shards_disjoint_and_sized(world):
    n = 60000
    parts = D.DataPartitioner(range(n), [1.0 / world] * world).partitions
    assert len(parts) == world
    k = int((1.0 / world) * n)
    assert all(len(p) == k for p in parts)
    flat = [i for p in parts for i in p]
    assert len(set(flat)) == len(flat)
    if world == 7:
        assert k == 8571 and n - len(flat) == 3


def test_same_seed_same_permutation_other_seed_differs():
    a = D.DataPartitioner(range(1000), [0.5, 0.5], seed=1234).partitions
    b = D.DataPartitioner(range(1000), [0.5, 0.5], seed=1234).partitions
    c = D.DataPartitioner(range(1000), [0.5, 0.5], seed=4321).partitions
    assert a == b and a != c


@settings(max_examples=40, deadline=None)
@given(n=st.integers(1, 500), world=st.integers(1, 9), seed=st.integers(0, 2 ** 31 - 1))
def test_partition_property(n, world, seed):
    parts = D.DataPartitioner(range(n), [1.0 / world] * world, seed=seed).partitions
    flat = [i for p in parts for i in p]
    assert len(flat) == len(set(flat)) and all(0 <= i < n for i in flat)
    assert all(len(p) == int((1.0 / world) * n) for p in parts)


def test_synthetic_dataset_deterministic_and_mnist_like():
    a, b = D.SyntheticMNIST(n=256, seed=7), D.SyntheticMNIST(n=256, seed=7)
    assert torch.equal(a.images, b.images) and torch.equal(a.labels, b.labels)
    x, y = a[5]
    assert x.shape == (1, 28, 28) and x.dtype == torch.float32 and isinstance(y, int) and 0 <= y < 10
    ref = (a.images[5].float() / 255.0 - D.MNIST_MEAN) / D.MNIST_STD
    assert torch.allclose(x[0], ref, atol=1e-6)


@pytest.mark.parametrize("world,bsz", [(1, 128), (2, 64), (4, 32), (8, 16), (7, 18)])
def test_partition_dataset_batch_rule(world, bsz):
    ds = D.SyntheticMNIST(n=1024, seed=1)
    seen = []
    for r in range(world):
        loader, b = D.partition_dataset(ds, rank=r, world_size=world, pin_memory=False)
        assert b == bsz == 128 // world
        assert len(loader.dataset) == int((1.0 / world) * 1024)
        seen += list(loader.dataset.index)
    assert len(seen) == len(set(seen))


def test_batch_loader_matches_per_sample_path():
    ds = D.SyntheticMNIST(n=300, seed=3)
    part = D.DataPartitioner(ds, [0.5, 0.5]).use(1)
    loader = D.BatchLoader(part, batch_size=32, shuffle=True, pin_memory=False, seed=11)
    assert len(loader) == 5
    total = 0
    seen = set()
    for x, y in loader:
        total += x.shape[0]
        assert x.shape[1:] == (1, 28, 28) and y.dtype == torch.int64
        # every row must be one of the partition's samples with the right label
        for row, lab in zip(x, y):
            hits = [i for i in part.index if int(ds.labels[i]) == int(lab)
                    and torch.allclose(ds[i][0], row, atol=1e-5)]
            assert hits
            seen.add(hits[0])
    assert total == 150 and len(seen) == 150


def test_batch_loader_raw_uint8():
    ds = D.SyntheticMNIST(n=64, seed=3)
    loader = D.BatchLoader(D.Partition(ds, list(range(64))), 16, shuffle=False, pin_memory=False,
                           raw_uint8=True)
    x, y = next(iter(loader))
    assert x.dtype == torch.uint8 and torch.equal(x[:, 0], ds.images[:16]) and torch.equal(y, ds.labels[:16])


def test_idx_roundtrip(tmp_path):
    ds = D.SyntheticMNIST(n=50, seed=9)
    D.write_idx(str(tmp_path), ds, test_n=10)
    back = D.load_mnist(str(tmp_path), train=True)
    assert torch.equal(back.images, ds.images) and torch.equal(back.labels, ds.labels)
    assert len(D.load_mnist(str(tmp_path), train=False)) == 10


def test_public_names():
    for name in ["init_processes", "init_process", "send", "recv", "isend", "irecv", "all_reduce", "reduce",
                 "broadcast", "scatter", "gather", "all_gather", "new_group", "reduce_op", "allreduce",
                 "average_gradients", "partition_dataset", "Partition", "DataPartitioner", "Net", "run",
                 "get_rank", "get_world_size", "launch"]:
        assert hasattr(b2, name), name
    assert {"SUM", "PRODUCT", "MAX", "MIN"} <= set(dir(b2.reduce_op))
