"""Worker functions for the distributed-CPU tier (run in spawned processes, gloo @127.0.0.1)."""
import torch
import torch.nn.functional as F

import dist_tuto.pth_b200 as b2
from dist_tuto.pth_b200 import ring
from dist_tuto.pth_b200.data import SyntheticMNIST
from dist_tuto.pth_b200.parallel.ddp import DistributedDataParallel, GradBucket


def w_p2p(rank, size):
    # blocking (tuto.md:82-91)
    t = torch.zeros(1)
    if rank == 0:
        t += 1
        b2.send(t, dst=1)
    elif rank == 1:
        b2.recv(t, src=0)
    if rank in (0, 1):
        assert t[0] == 1.0
    # non-blocking (tuto.md:102-116)
    t = torch.zeros(3)
    req = None
    if rank == 0:
        t += 7
        req = b2.isend(t, dst=1)
    elif rank == 1:
        req = b2.irecv(t, src=0)
    if req is not None:
        req.wait()
        assert req.is_completed()
        assert torch.equal(t, torch.full((3,), 7.0))
    b2.barrier()


def w_collectives(rank, size):
    ops = {"SUM": (b2.reduce_op.SUM, lambda v: sum(v)),
           "PRODUCT": (b2.reduce_op.PRODUCT, lambda v: torch.tensor(v).prod().item()),
           "MAX": (b2.reduce_op.MAX, max), "MIN": (b2.reduce_op.MIN, min)}
    vals = [float(r + 1) for r in range(size)]
    for name, (op, fn) in ops.items():
        t = torch.full((4,), float(rank + 1))
        b2.all_reduce(t, op=op)
        assert torch.allclose(t, torch.full((4,), float(fn(vals)))), name
        t = torch.full((4,), float(rank + 1))
        b2.reduce(t, dst=size - 1, op=op)
        if rank == size - 1:
            assert torch.allclose(t, torch.full((4,), float(fn(vals)))), name
    # group=0 means WORLD as in 2017 (train_dist.py:99)
    t = torch.ones(1)
    b2.all_reduce(t, op=b2.reduce_op.SUM, group=0)
    assert t[0] == size
    # broadcast
    t = torch.full((2,), float(rank))
    b2.broadcast(t, src=1)
    assert torch.equal(t, torch.ones(2))
    # scatter (list given on every rank, tutorial-era style)
    out = torch.zeros(2)
    lst = [torch.full((2,), float(10 + r)) for r in range(size)]
    b2.scatter(out, src=0, scatter_list=lst)
    assert torch.equal(out, torch.full((2,), float(10 + rank)))
    # gather as in ptp.py:21-28 (gather_list passed by every rank)
    tl = [torch.zeros(1) for _ in range(size)]
    b2.gather(torch.ones(1), dst=0, gather_list=tl, group=0)
    s = sum(tl)[0]
    assert s == (size if rank == 0 else 0)
    # the ptp.py:9-19 helper: gather(tensor, rank, tensor_list, root, group), non-zero root
    tl2 = [torch.zeros(1) for _ in range(size)]
    b2.gather_to_root(torch.full((1,), float(rank + 1)), rank, tl2 if rank == size - 1 else None, root=size - 1)
    if rank == size - 1:
        assert [float(t) for t in tl2] == [float(r + 1) for r in range(size)]
    # all_gather
    tl = [torch.zeros(1) for _ in range(size)]
    b2.all_gather(tl, torch.full((1,), float(rank)))
    assert [int(x[0]) for x in tl] == list(range(size))
    # new_group (tuto.md:180-185)
    g = b2.new_group([0, 1])
    t = torch.ones(1)
    if rank in (0, 1):
        b2.all_reduce(t, op=b2.reduce_op.SUM, group=g)
        assert t[0] == 2.0
    assert b2.get_world_size() == size and b2.get_rank() == rank
    b2.barrier()


def w_ring(rank, size):
    # rank-dependent data catches defect D3 (reference returns 3/6/9 instead of 6/6/6)
    send = torch.arange(6, dtype=torch.float32).view(2, 3) * (rank + 1)
    keep = send.clone()
    recv = torch.zeros(2, 3)
    b2.allreduce(send, recv)
    expect = torch.arange(6, dtype=torch.float32).view(2, 3) * sum(r + 1 for r in range(size))
    assert torch.allclose(recv, expect), (rank, recv)
    assert torch.equal(send, keep)                       # out-of-place (tuto.md:354)
    for n in (1, 5, 64, 1000):
        send = torch.randn(n, generator=torch.Generator().manual_seed(100 + rank))
        recv = torch.empty(n)
        ring.allreduce_chunked(send, recv)
        ref = sum(torch.randn(n, generator=torch.Generator().manual_seed(100 + r)) for r in range(size))
        assert torch.allclose(recv, ref, atol=1e-5), n
    if size >= 3:
        g = b2.new_group([0, 2])
        if rank in (0, 2):
            recv = torch.zeros(2)
            b2.allreduce(torch.full((2,), float(rank + 1)), recv, group=g)
            assert torch.equal(recv, torch.full((2,), 4.0))
    b2.barrier()


def _loss_grads(model, rank):
    g = torch.Generator().manual_seed(50 + rank)
    x = torch.randn(8, 1, 28, 28, generator=g)
    y = torch.randint(0, 10, (8,), generator=g)
    model.zero_grad()
    F.nll_loss(model(x), y).backward()


def w_average_gradients(rank, size):
    torch.manual_seed(1234)
    model = b2.Net().eval()                               # eval: no dropout, deterministic oracle
    _loss_grads(model, rank)
    local = [p.grad.clone() for p in model.parameters()]
    gathered = [[torch.zeros_like(g) for _ in range(size)] for g in local]
    for g, lst in zip(local, gathered):
        b2.all_gather(lst, g)
    b2.average_gradients(model)                           # catches defect D1 (no communication)
    for p, lst in zip(model.parameters(), gathered):
        mean = sum(lst) / size
        assert torch.allclose(p.grad, mean, atol=1e-6)
    # bucketed path
    _loss_grads(model, rank)
    model._grad_bucket = GradBucket(list(model.parameters()))
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(model.parameters(), model._grad_bucket.views))
    b2.average_gradients(model)
    for p, lst in zip(model.parameters(), gathered):
        assert torch.allclose(p.grad, sum(lst) / size, atol=1e-6)
    del model._grad_bucket
    # DDP wrapper, small cap -> several buckets; optimizer.zero_grad(set_to_none=True) tolerated
    torch.manual_seed(99 + rank)                          # different init per rank: broadcast must fix it
    m2 = b2.Net().eval()
    ddp = DistributedDataParallel(m2, bucket_cap_bytes=4096)
    assert len(ddp.buckets) > 2
    ref0 = [torch.zeros_like(p) for p in m2.parameters()]
    for p, r in zip(m2.parameters(), ref0):
        r.copy_(p.detach())
        b2.broadcast(r, src=0)
        assert torch.equal(r, p.detach())
    m3 = b2.Net().eval()
    m3.load_state_dict({k: v.clone() for k, v in m2.state_dict().items()})
    _loss_grads(m3, rank)
    gathered = []
    for p in m3.parameters():
        lst = [torch.zeros_like(p.grad) for _ in range(size)]
        b2.all_gather(lst, p.grad)
        gathered.append(lst)
    for it in range(2):
        if it == 0:
            ddp.zero_grad()
        else:
            for p in m2.parameters():
                p.grad = None
        g = torch.Generator().manual_seed(50 + rank)
        x = torch.randn(8, 1, 28, 28, generator=g)
        y = torch.randint(0, 10, (8,), generator=g)
        F.nll_loss(ddp(x), y).backward()
        b2.average_gradients(m2)
        for p, lst in zip(m2.parameters(), gathered):
            assert torch.allclose(p.grad, sum(lst) / size, atol=1e-6), it
    b2.barrier()


def w_flat_sgd_ddp(rank, size):
    """Tutorial loop with DistributedDataParallel + FlatSGD on N ranks == torch.optim.SGD on the global batch."""
    import torch.nn.functional as F
    torch.manual_seed(7)
    ref = b2.Net().eval()
    mine = b2.Net().eval()
    mine.load_state_dict({k: v.clone() for k, v in ref.state_dict().items()})
    ddp = DistributedDataParallel(mine, bucket_cap_bytes=16384)
    opt = b2.FlatSGD(ddp, lr=0.05, momentum=0.5)
    assert len(opt.buckets) > 1
    ref_opt = torch.optim.SGD(ref.parameters(), lr=0.05, momentum=0.5)
    for it in range(3):
        g = torch.Generator().manual_seed(200 + it)
        x, y = torch.randn(8 * size, 1, 28, 28, generator=g), torch.randint(0, 10, (8 * size,), generator=g)
        opt.zero_grad()
        F.nll_loss(ddp(x[rank * 8:(rank + 1) * 8]), y[rank * 8:(rank + 1) * 8]).backward()
        b2.average_gradients(ddp if it % 2 else mine)      # the wrapper and the wrapped module are both accepted
        opt.step()
        ref_opt.zero_grad()
        F.nll_loss(ref(x), y).backward()                   # mean over the global batch == mean of the per-rank means
        ref_opt.step()
    for (n, a), b in zip(ref.named_parameters(), mine.parameters()):
        assert torch.allclose(a, b, atol=1e-5), n
    for b in opt.buckets:                                  # step() left the buckets zeroed
        assert float(b.flat.abs().max()) == 0.0
    flat = torch.cat([p.detach().reshape(-1) for p in mine.parameters()])
    other = flat.clone()
    b2.broadcast(other, src=0)
    assert torch.equal(flat, other)                        # replicas identical
    ddp.remove_hooks()
    # bf16 gradients on the wire, fp32 master weights: same loop, looser tolerance, replicas still identical
    torch.manual_seed(7)
    ref2 = b2.Net().eval()
    mine2 = b2.Net().eval()
    mine2.load_state_dict({k: v.clone() for k, v in ref2.state_dict().items()})
    ddp2 = DistributedDataParallel(mine2, bucket_cap_bytes=16384, grad_dtype=torch.bfloat16)
    opt2 = b2.FlatSGD(ddp2, lr=0.05, momentum=0.5)
    ref_opt2 = torch.optim.SGD(ref2.parameters(), lr=0.05, momentum=0.5)
    for it in range(3):
        g = torch.Generator().manual_seed(300 + it)
        x, y = torch.randn(8 * size, 1, 28, 28, generator=g), torch.randint(0, 10, (8 * size,), generator=g)
        opt2.zero_grad()
        F.nll_loss(ddp2(x[rank * 8:(rank + 1) * 8]), y[rank * 8:(rank + 1) * 8]).backward()
        b2.average_gradients(mine2)
        opt2.step()
        ref_opt2.zero_grad()
        F.nll_loss(ref2(x), y).backward()
        ref_opt2.step()
    for (n, a), b in zip(ref2.named_parameters(), mine2.parameters()):
        assert torch.allclose(a, b, atol=3e-3, rtol=3e-2), n
    flat = torch.cat([p.detach().reshape(-1) for p in mine2.parameters()])
    other = flat.clone()
    b2.broadcast(other, src=0)
    assert torch.equal(flat, other)


def w_symm_fd_exchange(rank, size):
    """Host side of the symmetric-memory runtime without a GPU: the SCM_RIGHTS file-descriptor exchange over abstract
    unix datagram sockets (parallel/symm.py) delivers every rank's fd to every peer, twice in a row (tag check), and
    the one-sender form used for the multicast handle."""
    import os
    import socket
    import torch.distributed as dist
    from dist_tuto.pth_b200.parallel.symm import SymmWorld
    w = SymmWorld.__new__(SymmWorld)                        # plumbing only: no CUDA context, no allocations
    w.group, w.world, w.rank, w._tag = None, size, rank, 0
    tok = [os.urandom(6).hex() if rank == 0 else None]
    dist.broadcast_object_list(tok, src=0)
    w._token = tok[0]
    w._sock = socket.socket(socket.AF_UNIX, socket.SOCK_DGRAM)
    w._sock.bind(w._addr(rank))
    w._sock.settimeout(60.0)
    dist.barrier()
    for round_ in range(2):
        r_fd = os.memfd_create("b2-test")                   # shareable object, like a VMM allocation handle
        os.write(r_fd, b"from-%d-round-%d" % (rank, round_))
        got = w._exchange_fds(r_fd)                         # everyone sends its descriptor to everyone
        os.close(r_fd)
        assert sorted(got) == [r for r in range(size) if r != rank]
        for src, fd in got.items():
            assert os.pread(fd, 64, 0) == b"from-%d-round-%d" % (src, round_)  # the fd really is the sender's object
            os.close(fd)
    r_fd = os.memfd_create("b2-test-mc")
    os.write(r_fd, b"mc-handle")
    got = w._exchange_fds(r_fd, only_from=0)                # multicast-handle form: rank 0 -> all
    os.close(r_fd)
    if rank == 0:
        assert got == {}
    else:
        assert list(got) == [0] and os.pread(got[0], 64, 0) == b"mc-handle"
        os.close(got[0])
    w._sock.close()


def w_train(rank, size):
    ds = SyntheticMNIST(n=1024, seed=5)
    logs = []
    cfg = b2.TrainConfig(epochs=4, dataset=ds, engine="torch", device="cpu", lr=0.1,
                         log=lambda *a: logs.append(a))
    out = b2.train(rank, size, cfg)
    assert out["bsz"] == 128 // size and len(out["loss"]) == 4
    assert out["loss"][-1] < out["loss"][0] - 0.05
    assert logs[0][0] == "Rank " and logs[0][1] == rank
    # replicas stay bit-identical after training (sync SGD invariant)
    for p in out["model"].parameters():
        lst = [torch.zeros_like(p) for _ in range(size)]
        b2.all_gather(lst, p.detach())
        assert all(torch.equal(lst[0], t) for t in lst)
    b2.barrier()


def w_fail(rank, size):
    if rank == 1:
        raise ValueError("boom on rank 1")
    import time
    time.sleep(60)


def w_hang(rank, size):
    import time
    time.sleep(60)


def w_env(rank, size):
    assert b2.get_world_size() == size and b2.get_rank() == rank
    t = torch.ones(1)
    b2.all_reduce(t)
    assert t[0] == size


def w_one_node_guard(rank, size):
    """launch.assert_one_node(): passes on one host, raises a clear error when the ranks report different hosts."""
    import importlib
    import socket
    L = importlib.import_module("dist_tuto.pth_b200.launch")
    L.assert_one_node("b200")                              # same machine: fine
    real = socket.gethostname
    socket.gethostname = lambda: f"node-{rank}"            # pretend every rank sits on its own machine
    try:
        try:
            L.assert_one_node("b200")
        except RuntimeError as e:
            assert "ONE machine" in str(e) and "node-0" in str(e) and "nccl" in str(e)
        else:
            raise AssertionError("multi-host symmetric world was not rejected")
    finally:
        socket.gethostname = real


def w_aligned_start(rank, size):
    """bench_common.aligned_start: every rank of a single-node job leaves at the same CLOCK_MONOTONIC instant, whatever the
    skew with which the ranks arrive (rank 1 shows up 50 ms late)."""
    import os
    import sys
    import time
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench_common import aligned_start
    if rank == 1:
        time.sleep(0.05)
    t0 = aligned_start("cpu", lead_s=0.02)
    left = time.perf_counter()
    ts = [torch.zeros(2, dtype=torch.float64) for _ in range(size)]
    dist.all_gather(ts, torch.tensor([t0, left], dtype=torch.float64))
    starts = [float(t[0]) for t in ts]
    assert max(starts) - min(starts) < 1e-3, starts          # the agreed instant (spin exit: microseconds; CI noise: < 1 ms)
    assert all(float(t[1]) >= float(t[0]) for t in ts)


def w_hier_world(rank, size):
    """parallel/hier.py: 4 ranks pretending to be 2 machines x 2 ranks.  Level 1 (inside a machine), level 2 (one rail per
    local index, each reducing its slice across the machines) and the gather of the slices must add up to the global sum
    x scale for tiny, evenly divisible and ragged messages; launch._one_node() must see two machines."""
    import importlib
    import os
    import torch.distributed as dist
    L = importlib.import_module("dist_tuto.pth_b200.launch")
    from dist_tuto.pth_b200.parallel import hier, symm
    os.environ["B200DIST_FAKE_HOSTNAME"] = f"n{rank // 2}"
    try:
        assert L._one_node() is False
        w = hier.init_hier_world()
        assert symm.lookup_world(None) is w and hier.init_hier_world() is w
        d = w.describe()
        assert d["nodes"] == 2 and d["ranks_per_node"] == 2 and d["node"] == rank // 2 and d["local_rank"] == rank % 2
        assert w.nodes == [[0, 1], [2, 3]] and w.world == size
        for n in (10, 8192, 5001, 4097):
            g = torch.Generator().manual_seed(n)
            base = torch.randn(size, n, generator=g)                  # row r = rank r's contribution (same on every rank)
            t = base[rank].clone()
            out = w.all_reduce_(t, scale=1.0 / size)
            assert out.data_ptr() == t.data_ptr()
            want = base.sum(0) / size
            assert torch.allclose(t, want, atol=1e-5, rtol=1e-5), (n, float((t - want).abs().max()))
        # the CUDA branch (level 1 = a SymmWorld over the machine's ranks) with a stand-in that has SymmWorld's call shape
        class FakeLocal:
            multicast = False
            calls = []

            def __init__(self, group):
                self.group = group

            def supports(self, t):
                return t.is_contiguous()

            def all_reduce_(self, t, scale=1.0, handle=None, variant=None, wire=None, max_blocks=None):
                FakeLocal.calls.append((t.numel(), scale, handle, max_blocks))
                dist.all_reduce(t, group=self.group)
                t.mul_(scale)
                return t

            def alloc(self, numel, dtype):
                return ("symmetric", numel, dtype)

            def describe(self):
                return {"world": 2}

        w.local = FakeLocal(w.local_group)
        base = torch.arange(size * 6000, dtype=torch.float32).view(size, 6000)
        t = base[rank].clone()
        w.all_reduce_(t, scale=0.25, handle="H", max_blocks=7)
        assert torch.allclose(t, base.sum(0) * 0.25) and FakeLocal.calls == [(6000, 0.25, "H", 7)]
        assert w.alloc(10, torch.float32) == ("symmetric", 10, torch.float32) and w.describe()["local"] == {"world": 2}
        w.local = None
        hd = w.alloc(100, torch.float32)                              # plain bucket on the CPU
        assert hd.local.numel() == 128 and hd.local.dtype == torch.float32
        from dist_tuto.pth_b200.train import _spans_machines
        assert _spans_machines() is True
    finally:
        os.environ.pop("B200DIST_FAKE_HOSTNAME", None)
        symm._WORLDS.pop(None, None)
    assert L._one_node() is True                                      # real host names: one machine
    dist.barrier()


def w_train_trace(rank, size):
    """TrainConfig(trace=...): one Chrome trace with a process row per rank, epoch spans containing step-phase spans."""
    import json
    import os
    import tempfile
    path = os.path.join(tempfile.gettempdir(), f"b2_trace_{os.environ.get('MASTER_PORT', '0')}.json")
    out = b2.train(rank, size, b2.TrainConfig(epochs=2, dataset=SyntheticMNIST(n=512, seed=1), engine="torch", device="cpu",
                                              log=lambda *a: None, trace=path))
    assert (out["trace"] == path) == (rank == 0)
    if rank == 0:
        ev = json.load(open(path))["traceEvents"]
        os.remove(path)
        assert {e["pid"] for e in ev} == set(range(size))
        assert sorted(e["args"]["name"] for e in ev if e["ph"] == "M") == [f"rank {r}" for r in range(size)]
        for r in range(size):
            mine = [e for e in ev if e["pid"] == r and e["ph"] == "X"]
            epochs = [e for e in mine if e["cat"] == "epoch"]
            assert [e["name"] for e in epochs] == ["epoch 0", "epoch 1"] and epochs[0]["args"]["engine"] == "torch"
            nb = 512 // size // (128 // size)
            for name in ("h2d", "forward + loss", "backward", "average_gradients", "optimizer", "step"):
                spans = [e for e in mine if e["name"] == name]
                assert len(spans) == 2 * nb, (name, len(spans))
                for e in spans:                                  # every step phase lies inside one of the two epoch spans
                    assert any(ep["ts"] <= e["ts"] and e["ts"] + e["dur"] <= ep["ts"] + ep["dur"] + 1 for ep in epochs), name
            assert len([e for e in mine if e["name"] == "next batch"]) == 2 * (nb + 1)     # + the exhausted fetch of each epoch
