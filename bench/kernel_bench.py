"""Per-kernel device timings (CUDA events, L2 flushed between iterations) with roofline fractions.

Writes one JSON object to stdout.  Rooflines use MEASURED_PEAKS.json (copy bandwidth / cuBLAS bf16) when
present, else the profiling-guide fallback (6.65 TB/s, 1.59 PFLOP/s)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dist_tuto.pth_b200.ops import _ext  # noqa: E402
from dist_tuto.pth_b200.ops.convnet_fused import NPAR_ALLOC, pack_params  # noqa: E402
from dist_tuto.pth_b200.ops.gemm import linear_bf16  # noqa: E402
from dist_tuto.pth_b200.models.convnet import Net  # noqa: E402
from dist_tuto.pth_b200.utils.timers import l2_flush  # noqa: E402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d["hbm_gbs"] * 1e9, d["bf16_tflops"] * 1e12, "measured"
    return 6.65e12, 1.59e15, "fallback"


def timeit(fn, iters=20, warm=3, flush=True):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            l2_flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)     # us
    ts.sort()
    return {"us_median": ts[len(ts) // 2], "us_min": ts[0]}


def main():
    dev = torch.device("cuda", 0)
    C = _ext.C()
    hbm, flops, src = peaks()
    out = {"peaks": {"hbm_Bps": hbm, "bf16_flops": flops, "source": src}, "gemm": [], "convnet": []}
    shapes = [(128, 64, 320), (8192, 32, 256), (4096, 512, 1024), (8192, 4096, 4096), (16384, 1000, 512)]
    if "--big-only" in sys.argv:
        shapes = [(8192, 4096, 4096)]
    for M, N, K in shapes:
        a = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        w = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
        b = torch.randn(N, device=dev)
        t = timeit(lambda: linear_bf16(a, w, b, relu=True))
        tc = timeit(lambda: torch.relu(torch.nn.functional.linear(a, w, b.to(torch.bfloat16))))
        fl = 2.0 * M * N * K
        by = 2.0 * (M * K + N * K + M * N)
        roof_us = max(fl / flops, by / hbm) * 1e6
        out["gemm"].append({"M": M, "N": N, "K": K, **t, "cublas_us_median": tc["us_median"],
                            "tflops": fl / (t["us_median"] * 1e-6) / 1e12, "roofline_us": roof_us,
                            "frac_of_roofline": roof_us / t["us_median"], "of": src})
    if "--gemm-only" in sys.argv:
        print(json.dumps(out, indent=1))
        return
    torch.manual_seed(0)
    params = pack_params(Net().to(dev))
    for B in (16, 32, 64, 128, 256, 1024):
        x = torch.randn(B, 1, 28, 28, device=dev)
        y = torch.randint(0, 10, (B,), device=dev)
        grads = torch.zeros(NPAR_ALLOC, device=dev)
        acc = torch.zeros(2, device=dev)
        step = torch.zeros(1, dtype=torch.int64, device=dev)
        fb = timeit(lambda: C.convnet_step(params, grads, x, y, acc, None, None, step, 1, 0, True, 1.0 / B, 0.5, 0))
        fw = timeit(lambda: C.convnet_step(params, None, x, y, acc, None, None, step, 1, 0, False, 1.0 / B, 0.5, 0))
        mom = torch.zeros_like(params)
        dc = torch.zeros(1, dtype=torch.int32, device=dev)
        sg = timeit(lambda: C.allreduce_sgd([grads.data_ptr()], [0], params.clone(), mom, step, 0.01, 0.5, 1.0, 0, 1, True, 0, dc))
        flop = 2.9e6 * B
        row = {"B": B, "fwd_bwd": fb, "fwd": fw, "sgd": sg, "samples_per_s_kernel": B / (fb["us_median"] * 1e-6),
               "gflops": flop / (fb["us_median"] * 1e-6) / 1e9}
        for cl in (2, 4, 8):                    # one cluster of `cl` CTAs per sample (one wave only)
            if B * cl <= 128:
                row[f"fwd_bwd_cluster{cl}"] = timeit(lambda: C.convnet_step(params, grads, x, y, acc, None, None, step, 1, 0, True,
                                                                            1.0 / B, 0.5, 0, 0, cl))
        out["convnet"].append(row)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
