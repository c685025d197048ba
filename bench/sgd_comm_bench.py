"""Device time of the fused all-reduce + SGD kernel alone (CUDA-graph replay of 64 back-to-back calls), N ranks."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dist_tuto.pth_b200 as b2  # noqa: E402
from dist_tuto.pth_b200.ops.convnet_fused import FusedTrainer  # noqa: E402


def body(rank, size):
    dev = torch.device("cuda", torch.cuda.current_device())
    tr = FusedTrainer(128 // size, device=dev)
    st = tr.stream
    C = tr.C

    def sgd():
        C.allreduce_sgd(tr._grad_ptrs, tr._sig_ptrs, tr.params, tr.momentum, tr.step_counter, 0.0, 0.5, 1.0 / size, tr.rank,
                        size, True, tr.grad_stride, tr.done_counter, tr.aux)

    with torch.cuda.stream(st):
        for _ in range(3):
            sgd()
    st.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=st):
        for _ in range(64):
            sgd()
    res = []
    for _ in range(5):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            e0.record(st)
            gr.replay()
            e1.record(st)
        st.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) * 1e3 / 64], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res.append(float(t))
    if rank == 0:
        print(json.dumps({"n_gpus": size, "allreduce_sgd_us": round(min(res), 2), "pdl": os.environ.get("B200DIST_PDL", "1")}), flush=True)
    dist.barrier()


if __name__ == "__main__":
    b2.init_from_env(body, backend="b200")
