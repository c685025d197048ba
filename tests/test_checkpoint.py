"""Checkpoint / resume (absent in the reference, required of the framework)."""
import copy

import torch
import torch.nn.functional as F

import dist_tuto.pth_b200 as dist
from dist_tuto.pth_b200.utils.checkpoint import load_checkpoint, save_checkpoint


def _batch(i):
    g = torch.Generator().manual_seed(100 + i)
    return torch.randn(16, 1, 28, 28, generator=g), torch.randint(0, 10, (16,), generator=g)


def _step(model, opt, i):
    x, y = _batch(i)
    opt.zero_grad()
    F.nll_loss(model(x), y).backward()
    opt.step()


def test_module_and_flat_sgd_roundtrip_continues_exactly(tmp_path):
    torch.manual_seed(3)
    a = dist.Net().eval()
    oa = dist.FlatSGD(a, lr=0.05, momentum=0.9)
    for i in range(3):
        _step(a, oa, i)
    path = save_checkpoint(str(tmp_path / "ck.pt"), a, optimizer=oa, steps=3, history=[1.0])
    for i in range(3, 6):
        _step(a, oa, i)
    b = dist.Net().eval()                                   # different init: everything must come from the file
    ob = dist.FlatSGD(b, lr=0.01, momentum=0.1)
    blob = load_checkpoint(path, b, ob)
    assert blob["steps"] == 3 and blob["history"] == [1.0] and ob.momentum == 0.9 and ob.lr == 0.05
    for i in range(3, 6):
        _step(b, ob, i)
    for (n, p), q in zip(a.named_parameters(), b.parameters()):
        assert torch.equal(p, q), n                         # momentum restored -> bit-identical continuation
    # the parameters are still views of the optimizer's flat buffer after load_state_dict
    base, n = ob.param_flats[0].data_ptr(), ob.param_flats[0].numel()
    assert all(base <= p.data_ptr() < base + 4 * n for p in b.parameters())


def test_torch_optimizer_roundtrip(tmp_path):
    torch.manual_seed(4)
    a = dist.Net().eval()
    oa = torch.optim.SGD(a.parameters(), lr=0.05, momentum=0.5)
    _step(a, oa, 0)
    path = save_checkpoint(str(tmp_path / "sub" / "ck.pt"), a, optimizer=oa)      # creates the directory, atomic rename
    b = copy.deepcopy(a)
    for p in b.parameters():
        p.data.zero_()
    ob = torch.optim.SGD(b.parameters(), lr=0.05, momentum=0.5)
    load_checkpoint(path, b, ob)
    _step(a, oa, 1)
    _step(b, ob, 1)
    for p, q in zip(a.parameters(), b.parameters()):
        assert torch.equal(p, q)


def test_train_checkpoint_then_resume_accumulates_steps(tmp_path):
    ds = dist.SyntheticMNIST(n=1024, seed=5)
    ck1, ck2 = str(tmp_path / "a.pt"), str(tmp_path / "b.pt")
    quiet = lambda *a, **k: None  # noqa: E731
    out1 = dist.train(0, 1, dist.TrainConfig(epochs=1, max_steps=5, dataset=ds, engine="torch", device="cpu", checkpoint=ck1, log=quiet))
    assert out1["steps"] == 5
    blob1 = torch.load(ck1, map_location="cpu")
    assert blob1["steps"] == 5 and "optim" in blob1 and len(blob1["optim"]["momentum_buffers"]) == 1
    out2 = dist.train(0, 1, dist.TrainConfig(epochs=1, max_steps=3, dataset=ds, engine="torch", device="cpu", resume=ck1, checkpoint=ck2,
                                             log=quiet))
    blob2 = torch.load(ck2, map_location="cpu")
    assert out2["steps"] == 3 and blob2["steps"] == 8                      # cumulative across resumes
    # the resumed run started from the saved weights, not from the seed
    fresh = dist.Net()
    torch.manual_seed(1234)
    assert not torch.equal(blob2["model"]["fc2.bias"], fresh.state_dict()["fc2.bias"])
    assert float(blob2["optim"]["momentum_buffers"][0].abs().sum()) > 0


def test_named_momentum_roundtrip_across_layouts(tmp_path):
    """The per-name momentum written next to FlatSGD's own state restores into a fresh optimizer (the form the fused
    engine reads and writes), so checkpoints cross engines without losing momentum."""
    import torch.nn.functional as F
    from dist_tuto.pth_b200.utils.checkpoint import restore_optimizer
    torch.manual_seed(3)
    net = dist.Net().eval()
    opt = dist.FlatSGD(net, lr=0.05, momentum=0.5)
    x, y = torch.randn(8, 1, 28, 28), torch.randint(0, 10, (8,))
    for _ in range(2):
        opt.zero_grad()
        F.nll_loss(net(x), y).backward()
        opt.step()
    path = str(tmp_path / "ck.pt")
    save_checkpoint(path, net, optimizer=opt, steps=2)
    blob = torch.load(path)
    assert set(blob["momentum"]) == {n for n, _ in net.named_parameters()}
    assert float(blob["momentum"]["fc1.weight"].abs().max()) > 0
    # a consumer that only understands the named form (what FusedTrainer.load_state_dict reads)
    net2 = dist.Net().eval()
    net2.load_state_dict(blob["model"])
    opt2 = dist.FlatSGD(net2, lr=0.05, momentum=0.5)
    restore_optimizer(opt2, net2, {"momentum": blob["momentum"]})
    for a, b in zip(opt.momentum_flats, opt2.momentum_flats):
        assert torch.equal(a, b)
