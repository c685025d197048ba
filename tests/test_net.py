"""Unit tier (CPU): Net parity against an independent fp32 oracle written here."""
import torch
import torch.nn.functional as F

from dist_tuto.pth_b200.models.convnet import Net, PARAM_SHAPES, PARAM_NUMEL, param_offsets


def oracle(params, x):
    w1, b1, w2, b2_, w3, b3, w4, b4 = params
    h = F.conv2d(x, w1, b1)
    assert h.shape[1:] == (10, 24, 24)
    h = F.relu(F.max_pool2d(h, 2))            # pool BEFORE relu (train_dist.py:65)
    assert h.shape[1:] == (10, 12, 12)
    h = F.conv2d(h, w2, b2_)
    assert h.shape[1:] == (20, 8, 8)
    h = F.relu(F.max_pool2d(h, 2))
    assert h.shape[1:] == (20, 4, 4)
    h = h.reshape(-1, 320)
    h = F.relu(h @ w3.t() + b3)
    h = h @ w4.t() + b4
    return h - torch.logsumexp(h, dim=1, keepdim=True)


def test_param_inventory():
    net = Net()
    got = [(n, tuple(p.shape)) for n, p in net.named_parameters()]
    assert got == PARAM_SHAPES
    assert sum(p.numel() for p in net.parameters()) == PARAM_NUMEL == 21840
    offs, total = param_offsets()
    assert total == 21840 and offs["fc2.bias"] == 21830


def test_forward_matches_oracle_eval():
    torch.manual_seed(0)
    net = Net().eval()
    x = torch.randn(5, 1, 28, 28)
    out = net(x)
    ref = oracle([p.detach() for p in net.parameters()], x)
    assert out.shape == (5, 10)
    assert torch.allclose(out, ref, atol=1e-5)
    assert torch.allclose(out.exp().sum(1), torch.ones(5), atol=1e-5)


def test_same_seed_same_init():
    torch.manual_seed(1234)
    a = Net()
    torch.manual_seed(1234)
    b = Net()
    assert all(torch.equal(p, q) for p, q in zip(a.parameters(), b.parameters()))


def test_train_mode_has_dropout_eval_has_not():
    torch.manual_seed(0)
    net = Net()
    x = torch.randn(4, 1, 28, 28)
    net.train()
    assert not torch.allclose(net(x), net(x))
    net.eval()
    assert torch.allclose(net(x), net(x))
