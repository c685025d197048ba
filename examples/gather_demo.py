#!/usr/bin/env python
"""The reference's ``ptp.py`` scenario (ptp.py:21-28): every rank contributes ``ones(1)``, rank 0 gathers
them and prints the sum (= world size); other ranks print 0.

    python examples/gather_demo.py [--size 2] [--backend tcp|gloo|b200]

``--backend tcp`` (the reference default, ptp.py:30) is accepted and mapped to gloo."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dist_tuto.pth_b200 as dist  # noqa: E402
from dist_tuto.pth_b200.utils import say  # noqa: E402  (print as one write: ranks share the terminal)


def run(rank, size):
    say(dist.get_world_size())
    cuda = torch.cuda.is_available() and "nccl" in str(torch.distributed.get_backend())
    dev = torch.device("cuda", torch.cuda.current_device()) if cuda else torch.device("cpu")
    tensor = torch.ones(1, device=dev)
    tensor_list = [torch.zeros(1, device=dev) for _ in range(size)]
    dist.gather(tensor, dst=0, gather_list=tensor_list, group=0)     # group=0 == world, as in 2017
    say("Rank ", rank, " has data ", sum(tensor_list)[0].item())


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=2)
    ap.add_argument("--backend", default="tcp")
    a = ap.parse_args()
    dist.launch(run, size=a.size, backend=a.backend)
