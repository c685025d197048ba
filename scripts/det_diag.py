"""Where do two deterministic-mode runs of the fused trainer part ways?

Two ``FusedTrainer(deterministic=True)`` replicas are stepped in lock-step on identical batches; after every step the flat
parameters are compared bit for bit and, at the first mismatch, the per-CTA gradient slots of that step are compared slot by
slot and layer by layer (which CTA, which layer, how far apart).  Meant to be run plainly and under compute-sanitizer
(``scripts/runs/r2_c16.sh``): the bit-reproducibility test passes in plain runs and failed under memcheck with 4-CTA clusters.

    python scripts/det_diag.py --bsz 32 --steps 12 [--cluster 4] [--eager] [--out gpurun_out/det_diag.json]

Second part (``--tf32``): the multi-GPU "fused trainer == torch SGD on the global batch" check compares against a torch
model whose cuDNN convolutions default to TF32; this measures how far that reference itself moves with TF32 on / off.
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dist_tuto.pth_b200.ops.convnet_fused import LAYOUT, NPAR, NPAR_ALLOC, FusedTrainer, unpack_params  # noqa: E402

SEG = sorted(LAYOUT.items(), key=lambda kv: kv[1])


def seg_of(i):
    name = SEG[0][0]
    for n, o in SEG:
        if i >= o:
            name = n
    return name


def det_diag(args):
    dev = torch.device("cuda", 0)
    kw = dict(lr=0.05, seed=13, device=dev, p_drop=0.5, deterministic=True, use_graph=not args.eager)
    if args.cluster:
        kw["cluster"] = args.cluster
    a, b = FusedTrainer(args.bsz, **kw), FusedTrainer(args.bsz, **kw)
    g = torch.Generator().manual_seed(99)
    res = {"bsz": args.bsz, "cluster": a.cluster, "eager": args.eager, "steps": args.steps, "first_mismatch_step": None,
           "env": {k: os.environ.get(k) for k in ("B200DIST_PDL", "CUDA_LAUNCH_BLOCKING", "B200DIST_CONVNET_CLUSTER")}}
    for it in range(args.steps):
        x = torch.randn(args.bsz, 1, 28, 28, generator=g).pin_memory()
        y = torch.randint(0, 10, (args.bsz,), generator=g).pin_memory()
        for tr in (a, b):
            tr.step(x, y)
            tr.sync_lag(0)
        torch.cuda.synchronize()
        if not torch.equal(a.params, b.params):
            d = (a.params - b.params).abs()
            res["first_mismatch_step"] = it
            res["param_max_abs_diff"] = float(d.max())
            res["param_mismatch_by_layer"] = {}
            for n, t in unpack_params(d).items():
                res["param_mismatch_by_layer"][n] = {"n_diff": int((t != 0).sum()), "max": float(t.max())}
            pa = a.det_partials.view(-1, NPAR_ALLOC)[:args.bsz * a.cluster]
            pb = b.det_partials.view(-1, NPAR_ALLOC)[:args.bsz * b.cluster]
            sd = (pa - pb).abs()
            slots = []
            for s in torch.nonzero(sd.amax(dim=1) > 0).flatten().tolist():
                row = sd[s]
                idx = torch.nonzero(row > 0).flatten()
                layers = {}
                for i in idx.tolist():
                    layers.setdefault(seg_of(i) if i < NPAR else "loss_terms", []).append(i)
                slots.append({"slot": s, "sample": s // a.cluster, "cta_rank": s % a.cluster, "n_diff": int(idx.numel()),
                              "max": float(row.max()), "rel_to_slot_max": float(row.max() / (pa[s].abs().max() + 1e-30)),
                              "layers": {k: len(v) for k, v in layers.items()}})
            res["slots_differing"] = len(slots)
            res["slots"] = slots[:40]
            break
    print(json.dumps(res), flush=True)
    return res


def tf32_diag(args):
    """torch reference vs itself (fp32 vs TF32 convolutions) and vs the fused trainer, global batch ``gb``."""
    from dist_tuto.pth_b200.models.convnet import Net
    dev = torch.device("cuda", 0)
    out = {}
    for gb in (32, 80, 128):
        finals = {}
        for tf32 in (False, True):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.manual_seed(21)
            ref = Net(p_drop=0.0).to(dev)
            opt = torch.optim.SGD(ref.parameters(), lr=0.05, momentum=0.5)
            tr = FusedTrainer(gb, lr=0.05, momentum=0.5, seed=21, device=dev, p_drop=0.0, init_from=ref) if not tf32 else None
            for i in range(5):
                g = torch.Generator().manual_seed(300 + i)
                xg = torch.randn(gb, 1, 28, 28, generator=g)
                yg = torch.randint(0, 10, (gb,), generator=g)
                if tr is not None:
                    tr.step(xg.pin_memory(), yg.pin_memory())
                opt.zero_grad()
                F.nll_loss(ref(xg.to(dev)), yg.to(dev)).backward()
                opt.step()
            finals[tf32] = {n: p.detach().clone() for n, p in ref.named_parameters()}
            if tr is not None:
                tr.sync_lag(0)
                torch.cuda.synchronize()
                views = unpack_params(tr.params)
                out[f"gb{gb}_ours_vs_torch_fp32"] = max(float((views[n] - p).abs().max()) for n, p in finals[False].items())
        out[f"gb{gb}_torch_tf32_vs_torch_fp32"] = max(float((finals[True][n] - finals[False][n]).abs().max()) for n in finals[True])
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    print(json.dumps(out), flush=True)
    return out


def grads_diag(args):
    """Who is right at batch 128?  fused kernels (1 CTA / 4-CTA cluster per sample) and torch-GPU fp32 autograd, each
    against float64 autograd on the CPU (max abs gradient error over all parameters, one batch, no dropout)."""
    from dist_tuto.pth_b200.models.convnet import Net
    from dist_tuto.pth_b200.ops.convnet_fused import convnet_loss_and_grads, pack_params
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    out = {}
    for B in (32, 128):
        torch.manual_seed(5)
        net = Net(p_drop=0.0).eval()
        g = torch.Generator().manual_seed(17)
        x, y = torch.randn(B, 1, 28, 28, generator=g), torch.randint(0, 10, (B,), generator=g)
        n64 = Net(p_drop=0.0).double().eval()
        n64.load_state_dict({k: v.double() for k, v in net.state_dict().items()})
        F.nll_loss(n64(x.double()), y).backward()
        truth = {n: p.grad.float() for n, p in n64.named_parameters()}
        ngpu = Net(p_drop=0.0).to(dev).eval()
        ngpu.load_state_dict(net.state_dict())
        F.nll_loss(ngpu(x.to(dev)), y.to(dev)).backward()
        out[f"B{B}_torch_gpu_fp32"] = max(float((p.grad.cpu() - truth[n]).abs().max()) for n, p in ngpu.named_parameters())
        flat = pack_params(net, dev)
        for cl in ((1, 4) if B * 4 <= 148 else (1,)):
            _, gr = convnet_loss_and_grads(flat, x.to(dev), y.to(dev), training=False, cluster=cl)
            torch.cuda.synchronize()
            views = unpack_params(gr)
            out[f"B{B}_ours_cluster{cl}"] = max(float((views[n].cpu() - truth[n]).abs().max()) for n in truth)
        out[f"B{B}_grad_max"] = max(float(t.abs().max()) for t in truth.values())
    print(json.dumps(out), flush=True)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--bsz", type=int, default=32)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--cluster", type=int, default=0)
    ap.add_argument("--eager", action="store_true")
    ap.add_argument("--tf32", action="store_true")
    ap.add_argument("--grads", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    r = tf32_diag(args) if args.tf32 else (grads_diag(args) if args.grads else det_diag(args))
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        json.dump(r, open(args.out, "w"), indent=1)
