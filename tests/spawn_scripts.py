"""Tiny rank programs for tests/test_spawn.py (run as scripts by the launcher)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

mode = sys.argv[1]
rank = int(os.environ["RANK"])

if mode == "allreduce":
    import torch
    import dist_tuto.pth_b200 as dist

    def run(rank, size):
        t = torch.ones(1) * (rank + 1)
        dist.all_reduce(t)
        dist.utils.say("rank", rank, "of", size, "sum", t.item())

    dist.init_from_env(run, backend="gloo")
elif mode == "fail":
    if rank == 1:
        time.sleep(0.5)
        sys.exit(7)
    time.sleep(600)            # would hang forever without supervision
elif mode == "hang":
    time.sleep(600)
