"""Debug: native executor vs python loop vs eager loop, checksums after each epoch phase."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dist_tuto.pth_b200 import data as D
from dist_tuto.pth_b200.ops.convnet_fused import FusedTrainer
dev = torch.device("cuda:0")
ds = D.SyntheticMNIST(n=1000, seed=2)
part = D.Partition(ds, list(range(1000)))
def ck(tr):
    torch.cuda.synchronize()
    return f"{float(tr.params.double().sum()):.9f} {float(tr.params.double().abs().sum()):.9f} step={int(tr.step_counter.item())}"
for mode in ("native", "python", "eager"):
    loader = D.NativeBatchLoader(part, 64, seed=9, raw_uint8=True, pin_memory=True, num_buffers=int(os.environ.get("DBG_NB", "6")))
    tr = FusedTrainer(64, lr=0.05, seed=3, device=dev, p_drop=0.5, raw_uint8=True, use_graph=(mode != "eager"))
    out = []
    for ep, budget in ((0, None), (1, 6), (2, None)):
        if mode == "native":
            tr.run_native(loader, max_steps=budget)
        else:
            for i, (x, y) in enumerate(loader):
                if budget is not None and i == budget:
                    break
                tr.step(x, y)
        out.append(ck(tr))
    print(mode, *out, sep="\n   ")
