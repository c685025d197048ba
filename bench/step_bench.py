"""Warm (CUDA-graph replay) time of one fused training step on ONE GPU for the per-GPU batch sizes that occur at
1/2/4/8 GPUs (128/64/32/16 samples) and every cluster size -- the device-side cost of a step without the cross-GPU part."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dist_tuto.pth_b200.ops.convnet_fused import FusedTrainer  # noqa: E402

dev = torch.device("cuda", 0)
out = []
for bsz in (128, 64, 32, 16):
    for cl in (1, 2, 4, 8):
        if bsz * cl > 128 and cl > 1:
            continue
        tr = FusedTrainer(bsz, device=dev, cluster=cl, p_drop=0.5)
        pool = 64
        g = torch.Generator(device=dev).manual_seed(1)
        px = torch.randn(pool, bsz, 1, 28, 28, device=dev, generator=g)
        py = torch.randint(0, 10, (pool, bsz), device=dev, generator=g)
        st = tr.stream
        with torch.cuda.stream(st):
            for i in range(5):
                tr._kernels(px[i], py[i], bsz)
        st.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st):
            for i in range(pool):
                tr._kernels(px[i], py[i], bsz)
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(st):
                e0.record(st)
                gr.replay()
                e1.record(st)
            st.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3 / pool)
        out.append({"bsz": bsz, "cluster": cl, "us_per_step": round(best, 2), "samples_per_s": round(bsz / (best * 1e-6))})
        print(json.dumps(out[-1]), flush=True)
json.dump(out, open(os.environ.get("B2_OUT", "gpurun_out/step_bench.json"), "w"), indent=1)
