#!/usr/bin/env python
"""Launches 2 (or N) ranks that run a few fused all-reduces of each variant + fused trainer steps; meant to be profiled as a
WHOLE with `ncu --replay-mode application --target-processes all` (kernel replay cannot be used on kernels that spin on
peers: the peers would not replay)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dist_tuto.pth_b200 as b2  # noqa: E402
from dist_tuto.pth_b200.parallel import symm  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
NBYTES = int(sys.argv[2]) if len(sys.argv) > 2 else (64 << 20)


def body(rank, size):
    dev = torch.device("cuda", torch.cuda.current_device())
    w = symm.lookup_world(None)
    n = NBYTES // 4
    hd = w.alloc(n, torch.float32)
    hd.local.fill_(1.0)
    small = w.alloc(21888, torch.float32)
    small.local.fill_(1.0)
    torch.cuda.synchronize()
    b2.barrier()
    for v in [0, 1] + ([2] if w.multicast else []):
        for _ in range(2):
            if v != 0:
                w.all_reduce_(hd.local, scale=1.0 / size, handle=hd, variant=v)
            w.all_reduce_(small.local, scale=1.0 / size, handle=small, variant=v)
    torch.cuda.synchronize()
    from dist_tuto.pth_b200.ops.convnet_fused import FusedTrainer
    bsz = 128 // size
    tr = FusedTrainer(bsz, seed=1, device=dev)
    g = torch.Generator(device=dev).manual_seed(rank)
    x = torch.randn(bsz, 1, 28, 28, device=dev, generator=g)
    y = torch.randint(0, 10, (bsz,), device=dev, generator=g)
    with torch.cuda.stream(tr.stream):
        for _ in range(4):
            tr._kernels(x, y, bsz)
    tr.stream.synchronize()
    b2.barrier()


if __name__ == "__main__":
    b2.launch(body, size=N, backend="b200", join_timeout_s=600)
