"""ResNet-18 bf16 data-parallel step throughput (BASELINE.json config #3).

ours      : ResNet18 (channels_last, bf16 autocast) + DistributedDataParallel (symmetric flat buckets, fused
            one/two-shot/NVLS all-reduce with the 1/N scale, overlapped with backward on a side stream) + SGD.
reference : same model + the tutorial's per-parameter ``dist.all_reduce`` + ``/= size`` after backward (62
            NCCL calls + 62 divides per step, no overlap) -- tuto.md:310-314 semantics.
torch_ddp : same model in ``torch.nn.parallel.DistributedDataParallel`` (NCCL, 25 MB buckets overlapped with backward,
            ``gradient_as_bucket_view``) -- the library's own answer to the tutorial's closing advice (tuto.md:320), i.e. the
            honest competitor of the bucketed path.
nocomm    : the same step with NO gradient exchange at all: (arm - nocomm) is the communication a step still exposes.
ours_bf16 : our arm with bf16 gradient buckets on the wire (fp32 accumulation in the kernel, fp32 master weights).
Device-timed (CUDA events), max over ranks; images/s whole job.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dist_tuto.pth_b200 as b2  # noqa: E402
from dist_tuto.pth_b200.models.resnet import ResNet18  # noqa: E402
from dist_tuto.pth_b200.parallel.ddp import DistributedDataParallel  # noqa: E402

import types
ARGS = types.SimpleNamespace(**json.loads(os.environ["B2_BENCH_ARGS"])) if "B2_BENCH_ARGS" in os.environ else None


def run_mode(mode, rank, size, dev):
    torch.manual_seed(1234)
    model = ResNet18(num_classes=1000).to(dev).to(memory_format=torch.channels_last)
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.5)
    ddp, tddp = None, None
    if mode in ("ours", "ours_bf16"):
        ddp = DistributedDataParallel(model, bucket_cap_bytes=ARGS.bucket_mb << 20,
                                      **({"grad_dtype": torch.bfloat16} if mode == "ours_bf16" else {}))
        if getattr(ARGS, "flat_sgd", True):     # one sgd_flat launch per bucket (update + re-zero), params become views
            opt = b2.FlatSGD(ddp, lr=0.01, momentum=0.5)
    elif mode == "torch_ddp":
        tddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index], gradient_as_bucket_view=True)
    B = ARGS.batch
    x = torch.randn(B, 3, ARGS.res, ARGS.res, device=dev).to(memory_format=torch.channels_last)
    y = torch.randint(0, 1000, (B,), device=dev)

    def step():
        if ddp is not None:
            opt.zero_grad() if isinstance(opt, b2.FlatSGD) else ddp.zero_grad()
        else:
            opt.zero_grad(set_to_none=False)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = F.cross_entropy((tddp if tddp is not None else model)(x), y)
        loss.backward()
        if ddp is not None:
            b2.average_gradients(model)
        elif mode in ("torch_ddp", "nocomm"):
            pass                                   # torch DDP reduced inside backward / no exchange at all
        else:
            n = float(size)
            for p in model.parameters():
                dist.all_reduce(p.grad.data, op=dist.ReduceOp.SUM)
                p.grad.data /= n
        opt.step()
        return loss

    for _ in range(ARGS.warmup):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(ARGS.steps):
        loss = step()
    e1.record()
    e1.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / ARGS.steps
    if ddp is not None:
        ddp.remove_hooks()
    return {"ms_per_step": ms, "images_per_s": B * size / (ms * 1e-3), "loss": float(loss)}


def body(rank, size):
    dev = torch.device("cuda", torch.cuda.current_device())
    torch.backends.cudnn.benchmark = True
    res = {"n_gpus": size, "per_gpu_batch": ARGS.batch, "res": ARGS.res, "bucket_mb": ARGS.bucket_mb}
    for mode in ("nocomm", "reference", "torch_ddp", "ours", "ours_bf16"):
        res[mode] = run_mode(mode, rank, size, dev)
        torch.cuda.empty_cache()
    base = res["nocomm"]["ms_per_step"]
    for mode in ("reference", "torch_ddp", "ours", "ours_bf16"):
        res[mode]["exposed_comm_ms"] = res[mode]["ms_per_step"] - base
    res["speedup"] = res["reference"]["ms_per_step"] / res["ours"]["ms_per_step"]
    res["vs_torch_ddp"] = res["torch_ddp"]["ms_per_step"] / res["ours"]["ms_per_step"]
    res["vs_torch_ddp_bf16_wire"] = res["torch_ddp"]["ms_per_step"] / res["ours_bf16"]["ms_per_step"]
    if rank == 0:
        print(json.dumps(res), flush=True)
        os.makedirs(os.path.dirname(ARGS.out) or ".", exist_ok=True)
        json.dump(res, open(ARGS.out, "w"), indent=1)
    dist.barrier()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--res", type=int, default=224)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--bucket-mb", type=int, default=8)
    ap.add_argument("--no-flat-sgd", dest="flat_sgd", action="store_false", help="torch.optim.SGD instead of FlatSGD on our arm")
    ap.add_argument("--out", default=None)
    ARGS = ap.parse_args()
    ARGS.out = ARGS.out or f"gpurun_out/resnet_{ARGS.gpus}.json"
    os.environ["B2_BENCH_ARGS"] = json.dumps(vars(ARGS))
    if "RANK" in os.environ:
        b2.init_from_env(body, backend="b200")
    else:
        b2.launch(body, size=ARGS.gpus, backend="b200", join_timeout_s=1500)
