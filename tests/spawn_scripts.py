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
elif mode == "train_restart":
    # 3 epochs with a checkpoint after every epoch; on the FIRST attempt rank 0 (the rank that writes the checkpoints) dies at
    # the end of epoch 1, before that epoch's checkpoint -> the launcher starts the job over, which resumes from the checkpoint
    # written after epoch 0 and does the remaining two epochs
    import torch
    import dist_tuto.pth_b200 as dist
    from dist_tuto.pth_b200.data import SyntheticMNIST
    ckpt = sys.argv[2]
    attempt = int(os.environ.get("B200DIST_RESTART_COUNT", "0"))

    def run(rank, size):
        seen = []

        def log(*a):
            epoch = a[3]
            seen.append(epoch)
            if attempt == 0 and rank == 0 and epoch == 1:
                os._exit(9)

        cfg = dist.TrainConfig(epochs=3, dataset=SyntheticMNIST(n=512, seed=3), lr=0.05, engine="torch", device="cpu", log=log,
                               checkpoint=ckpt, checkpoint_every=1,
                               resume=ckpt if (attempt > 0 and os.path.exists(ckpt)) else None)
        out = dist.train(rank, size, cfg)
        if rank == 0:
            with open(ckpt + ".epochs", "a") as f:
                f.write(f"attempt {attempt}: epochs {seen} history {len(out['loss'])}\n")

    dist.init_from_env(run, backend="gloo")
