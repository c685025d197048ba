#!/usr/bin/env python
"""Collectives on a sub-group (tuto.md:176-186): all-reduce ``ones(1)`` over ranks {0, 1} -> 2.0 on both."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dist_tuto.pth_b200 as dist  # noqa: E402
from dist_tuto.pth_b200.utils import say  # noqa: E402  (print as one write: ranks share the terminal)


def run(rank, size):
    group = dist.new_group([0, 1])
    tensor = torch.ones(1)
    if rank in (0, 1):
        dist.all_reduce(tensor, op=dist.reduce_op.SUM, group=group)
    say("Rank ", rank, " has data ", tensor[0].item())


if __name__ == "__main__":
    dist.launch(run, size=int(sys.argv[1]) if len(sys.argv) > 1 else 2, backend="gloo")
