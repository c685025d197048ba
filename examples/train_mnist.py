#!/usr/bin/env python
"""Distributed synchronous SGD on (synthetic) MNIST -- the reference's ``train_dist.py`` scenario.

    python examples/train_mnist.py                       # CPU, gloo, world 2 (the reference default)
    python examples/train_mnist.py --backend b200 --size 8     # one process per B200, fused engine
    python examples/train_mnist.py --backend b200 --size 8 --global-batch 32768   # large batch: tcgen05 batched engine
    python -m dist_tuto.pth_b200.spawn --size 8 --max-restarts 2 examples/train_mnist.py --external --backend b200 \
        --checkpoint run.pt --checkpoint-every 1       # supervised: a failed job is restarted and resumes from run.pt
    torchrun --nproc-per-node 8 examples/train_mnist.py --backend b200 --external

Prints ``Rank r, epoch e: mean loss`` per epoch like train_dist.py:125-127."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dist_tuto.pth_b200 as dist  # noqa: E402

import json  # noqa: E402


def run(rank, size):
    CFG = json.loads(os.environ["B2_TRAIN_CFG"])          # spawned ranks re-import this file: pass the CLI through the env
    resume = CFG["resume"]
    if resume is None and CFG["ckpt"] and int(os.environ.get("B200DIST_RESTART_COUNT", "0")) > 0 and os.path.exists(CFG["ckpt"]):
        resume = CFG["ckpt"]                              # restarted by the launcher: continue from our own last checkpoint
    cfg = dist.TrainConfig(epochs=CFG["epochs"], lr=CFG["lr"], max_steps=CFG["max_steps"], checkpoint=CFG["ckpt"],
                           checkpoint_every=CFG["ckpt_every"], resume=resume, global_batch=CFG["global_batch"],
                           engine=CFG["engine"], trace=CFG["trace"])
    out = dist.train(rank, size, cfg)
    if rank == 0:
        print(f"{out['steps']} steps, {out['samples_per_s']:.0f} samples/s (wall clock, whole job)")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=2)
    ap.add_argument("--backend", default="gloo")
    ap.add_argument("--epochs", type=int, default=10)
    ap.add_argument("--lr", type=float, default=0.01)
    ap.add_argument("--max-steps", type=int, default=None)
    ap.add_argument("--checkpoint", default=None)
    ap.add_argument("--checkpoint-every", type=int, default=None, help="also checkpoint after every N-th epoch")
    ap.add_argument("--resume", default=None)
    ap.add_argument("--trace", default=None, help="write a Chrome / Perfetto trace of all ranks to this file")
    ap.add_argument("--global-batch", type=int, default=128, help="split over the ranks (train_dist.py:85: 128 // world)")
    ap.add_argument("--engine", default="auto", choices=["auto", "torch", "fused", "batched"])
    ap.add_argument("--external", action="store_true", help="rank/size from torchrun/mpirun env")
    a = ap.parse_args()
    os.environ["B2_TRAIN_CFG"] = json.dumps(dict(epochs=a.epochs, lr=a.lr, max_steps=a.max_steps, ckpt=a.checkpoint, ckpt_every=a.checkpoint_every, resume=a.resume, trace=a.trace,
                                                 global_batch=a.global_batch, engine=a.engine))
    if a.external:
        dist.init_from_env(run, backend=a.backend)
    else:
        dist.launch(run, size=a.size, backend=a.backend)
