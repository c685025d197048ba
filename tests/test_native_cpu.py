"""CPU tier for the native runtime: the extension must import without a GPU driver, the C++ prefetcher must
reproduce the Python loader's contract, and the layout constants must agree between Python and CUDA."""
import os

import pytest
import torch

from dist_tuto.pth_b200 import data as D
from dist_tuto.pth_b200.ops import _ext

pytestmark = pytest.mark.skipif(not os.path.isfile(_ext.so_path()), reason="native extension not built")


def test_extension_imports_on_cpu_and_layout_matches():
    C = _ext.C()
    from dist_tuto.pth_b200.ops.convnet_fused import LAYOUT, NPAR, NPAR_ALLOC, unpack_params
    from dist_tuto.pth_b200.models.convnet import PARAM_SHAPES
    assert C.convnet_npar() == NPAR == 21848 and NPAR_ALLOC % 64 == 0
    assert C.convnet_smem_bytes() <= 227 * 1024
    flat = torch.arange(NPAR_ALLOC, dtype=torch.float32)
    views = unpack_params(flat)
    end = 0
    for name, shape in PARAM_SHAPES:
        assert LAYOUT[name] % 4 == 0 and LAYOUT[name] >= end          # 16-byte aligned, non-overlapping
        assert tuple(views[name].shape) == shape and float(views[name].flatten()[0]) == LAYOUT[name]
        end = LAYOUT[name] + views[name].numel()
    assert end <= NPAR


def test_pack_unpack_roundtrip():
    from dist_tuto.pth_b200.models.convnet import Net
    from dist_tuto.pth_b200.ops.convnet_fused import pack_params, unpack_params
    net = Net()
    flat = pack_params(net)
    for (name, p) in net.named_parameters():
        assert torch.equal(unpack_params(flat)[name], p.detach())
    assert float(flat[250:252].abs().sum()) == 0.0                     # padding


@pytest.mark.parametrize("raw", [False, True])
def test_native_loader_matches_dataset(raw):
    ds = D.SyntheticMNIST(n=300, seed=3)
    part = D.DataPartitioner(ds, [0.5, 0.5]).use(0)
    loader = D.NativeBatchLoader(part, 32, pin_memory=False, seed=7, raw_uint8=raw)
    assert len(loader) == 5 and len(loader.dataset) == 150
    orders = []
    for _ in range(2):
        seen, n = [], 0
        for x, y in loader:
            n += x.shape[0]
            assert x.shape[1:] == (1, 28, 28) and x.dtype == (torch.uint8 if raw else torch.float32)
            for row, lab in zip(x, y):
                if raw:
                    hits = [i for i in part.index if int(ds.labels[i]) == int(lab) and torch.equal(ds.images[i], row[0])]
                else:
                    hits = [i for i in part.index if int(ds.labels[i]) == int(lab) and torch.allclose(ds[i][0], row, atol=1e-5)]
                assert hits
                seen.append(hits[0])
        assert n == 150 and len(set(seen)) == 150
        orders.append(seen)
    assert orders[0] != orders[1]                                       # reshuffled every epoch


def test_partition_dataset_native_flag():
    ds = D.SyntheticMNIST(n=256, seed=1)
    l1, b1 = D.partition_dataset(ds, rank=0, world_size=2, native=True, pin_memory=False)
    l2, b2_ = D.partition_dataset(ds, rank=0, world_size=2, native=False, pin_memory=False)
    assert b1 == b2_ == 64 and len(l1) == len(l2) == 2
    assert isinstance(l1, D.NativeBatchLoader) and isinstance(l2, D.BatchLoader)


def test_gemm_reports_unavailable_without_driver():
    C = _ext.C()
    if not torch.cuda.is_available():
        assert C.gemm_available() is False
