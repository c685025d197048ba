#!/usr/bin/env python
"""Large-batch throughput of the batched tensor-core engine vs the per-sample SIMT engine (1 GPU).

    python bench/batched_bench.py --batch 1024 4096 16384 --steps 30 [--out gpurun_out/batched_bench.json]

Per batch size: full training step (forward + loss + backward + fused SGD + weight re-pack) as a CUDA graph, timed with
CUDA events over distinct device-resident batches (pool > L2 or L2 flushed, stated in the output); per-kernel times of
the batched engine (each stage timed alone on valid inputs); the per-sample SIMT engine on the same batch for reference.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dist_tuto.pth_b200.ops import _ext  # noqa: E402
from dist_tuto.pth_b200.ops.convnet_batched import STAGES, BatchedTrainer  # noqa: E402
from dist_tuto.pth_b200.ops.convnet_fused import FusedTrainer  # noqa: E402


def timed(fn, iters, stream):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(iters):
            fn()
        e1.record(stream)
    stream.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3      # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, nargs="+", default=[1024, 4096, 16384])
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "batched_bench.json"))
    ap.add_argument("--no-simt", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    C = _ext.C()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = {"gpu": torch.cuda.get_device_name(0), "steps": args.steps, "runs": []}
    for B in args.batch:
        g = torch.Generator(device=dev).manual_seed(B)
        npool = max(2, min(args.steps, (512 << 20) // (B * 784)))
        xs = torch.randint(0, 256, (npool, B, 1, 28, 28), dtype=torch.uint8, device=dev, generator=g)
        ys = torch.randint(0, 10, (npool, B), device=dev, generator=g)
        run = {"B": B, "pool_batches": npool, "pool_mb": npool * B * 784 / 2**20}
        # ------------------------------------------------ batched tensor-core engine
        tr = BatchedTrainer(B, seed=1, device=dev, raw_uint8=True)
        st = tr.stream
        with torch.cuda.stream(st):
            for i in range(3):
                tr._kernels(xs[i % npool], ys[i % npool], B)
        st.synchronize()
        graphs = []
        for i in range(npool):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=st):
                tr._kernels(xs[i], ys[i], B)
            graphs.append(gr)
        with torch.cuda.stream(st):
            for gr in graphs:
                gr.replay()
            flush.fill_(0)
        st.synchronize()
        it = {"i": 0}

        def one():
            graphs[it["i"] % npool].replay()
            it["i"] += 1
        us = timed(one, args.steps, st)
        run["batched_us_per_step"] = us
        run["batched_samples_per_s"] = B / us * 1e6
        run["loss_finite"] = bool(torch.isfinite(tr.loss_acc).all())
        # per-kernel: each stage alone (inputs of that stage are valid from the full steps above)
        stages = {}
        bl = tr.bufs.as_list()
        for name, bit in STAGES.items():
            if name == "all":
                continue
            def stage(bit=bit):
                C.bt_step(tr.params, tr.grads, xs[0], ys[0], bl, tr.loss_acc, None, tr.step_counter, 1, 0, True, 1.0 / B, 0.5, bit)
            with torch.cuda.stream(st):
                stage()
            st.synchronize()
            stages[name] = timed(stage, 10, st)
        def opt():
            C.allreduce_sgd(tr._grad_ptrs, tr._sig_ptrs, tr.params, tr.momentum, tr.step_counter, 0.0, 0.5, 1.0, 0, 1, True, 0,
                            tr.done_counter, None, [])
            C.bt_pack_weights(tr.params, bl)
        stages["sgd+pack"] = timed(opt, 10, st)
        run["stage_us"] = stages
        del tr, graphs
        # ------------------------------------------------ per-sample SIMT engine on the same batch
        if not args.no_simt:
            ft = FusedTrainer(B, seed=1, device=dev, raw_uint8=True, cluster=1)
            st2 = ft.stream
            with torch.cuda.stream(st2):
                for i in range(3):
                    ft._kernels(xs[i % npool], ys[i % npool], B)
            st2.synchronize()
            g2 = []
            for i in range(min(npool, 8)):
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=st2):
                    ft._kernels(xs[i], ys[i], B)
                g2.append(gr)
            with torch.cuda.stream(st2):
                for gr in g2:
                    gr.replay()
            st2.synchronize()
            j = {"i": 0}

            def one2():
                g2[j["i"] % len(g2)].replay()
                j["i"] += 1
            us2 = timed(one2, max(4, args.steps // 3), st2)
            run["simt_us_per_step"] = us2
            run["simt_samples_per_s"] = B / us2 * 1e6
            run["speedup"] = us2 / us
            del ft, g2
        res["runs"].append(run)
        print(json.dumps(run), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
