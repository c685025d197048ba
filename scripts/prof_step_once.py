#!/usr/bin/env python
"""A handful of training steps of one engine on cuda:0 -- the target of `ncu` captures (scripts/prof_*.sh).

    python scripts/prof_step_once.py batched 4096     # batched tensor-core engine, per-GPU batch 4096
    python scripts/prof_step_once.py fused 128        # per-sample fused engine (global batch 128 on one GPU)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dist_tuto.pth_b200.ops.convnet_batched import BatchedTrainer  # noqa: E402
from dist_tuto.pth_b200.ops.convnet_fused import FusedTrainer  # noqa: E402

engine = sys.argv[1] if len(sys.argv) > 1 else "batched"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
tr = (BatchedTrainer if engine == "batched" else FusedTrainer)(B, seed=1, device=dev, raw_uint8=True)
g = torch.Generator(device=dev).manual_seed(0)
xs = torch.randint(0, 256, (4, B, 1, 28, 28), dtype=torch.uint8, device=dev, generator=g)
ys = torch.randint(0, 10, (4, B), device=dev, generator=g)
with torch.cuda.stream(tr.stream):
    for i in range(6):
        tr._kernels(xs[i % 4], ys[i % 4], B)
tr.stream.synchronize()
print("done", float(tr.loss_acc[0]))
