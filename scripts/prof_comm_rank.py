#!/usr/bin/env python
"""ONE rank of a small comm workload (RANK / WORLD_SIZE / MASTER_PORT from the environment).  scripts/prof_comm.sh starts
rank 0 under `ncu` (single-pass metrics only: a kernel that spins on its peers cannot be replayed) and the other ranks plain."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dist_tuto.pth_b200 as b2  # noqa: E402
from dist_tuto.pth_b200.parallel import symm  # noqa: E402

NBYTES = int(os.environ.get("PROF_BYTES", 64 << 20))


def body(rank, size):
    dev = torch.device("cuda", torch.cuda.current_device())
    w = symm.lookup_world(None)
    hd = w.alloc(NBYTES // 4, torch.float32)
    hd.local.fill_(1.0)
    small = w.alloc(21888, torch.float32)
    small.local.fill_(1.0)
    torch.cuda.synchronize()
    b2.barrier()
    for v in [3, 0, 1] + ([2] if w.multicast else []):
        for _ in range(2):
            if v in (1, 2):
                w.all_reduce_(hd.local, scale=1.0 / size, handle=hd, variant=v)
            w.all_reduce_(small.local, scale=1.0 / size, handle=small, variant=v)
            torch.cuda.synchronize()
            b2.barrier()
    from dist_tuto.pth_b200.ops.convnet_fused import FusedTrainer
    bsz = 128 // size
    tr = FusedTrainer(bsz, seed=1, device=dev)
    g = torch.Generator(device=dev).manual_seed(rank)
    x = torch.randn(bsz, 1, 28, 28, device=dev, generator=g)
    y = torch.randint(0, 10, (bsz,), device=dev, generator=g)
    for _ in range(4):
        with torch.cuda.stream(tr.stream):
            tr._kernels(x, y, bsz)
        tr.stream.synchronize()
        b2.barrier()
    print("rank", rank, "done", flush=True)


if __name__ == "__main__":
    b2.init_from_env(body, backend="b200")
