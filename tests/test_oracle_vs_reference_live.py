"""CPU, this container only (skipped where /root/reference does not exist, e.g. on the GPU box): the float half of the oracle
against the UNMODIFIED reference modules on random shapes and hyper-parameters — the committed fixtures pin five fixed cases
per layer type, this sweeps beyond them (non-square maps, dilation, stride 1-3, 1x1 to 5x5 filters, batch 1, bias on/off).
Noise is replayed as SURVEY §8c describes: seed -> reference forward; same seed -> draw the same tensors in the same order."""
import numpy as np
import pytest
import torch

import oracle.qbn_oracle as O
from oracle.ref_harness import import_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="the reference tree is only mounted in the build container")


@pytest.fixture(scope="module")
def ref():
    import_reference()
    from src.models.stochastic.bbb.conv import Conv2d
    from src.models.stochastic.bbb.linear import Linear
    from src.models.stochastic.mcdropout.dropout import BernoulliDropout
    return Conv2d, Linear, BernoulliDropout


def _trained_like(mod, g):
    with torch.no_grad():
        mod.weight.copy_(torch.randn(mod.weight.shape, generator=g) / mod.weight[0].numel() ** 0.5)
        mod.std.copy_(torch.empty(mod.std.shape).uniform_(-5.0, -1.0, generator=g))
        if mod.bias is not None:
            mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.1)


def close(a, b, rtol=1e-5, atol=1e-6):
    np.testing.assert_allclose(np.asarray(a.detach() if torch.is_tensor(a) else a), np.asarray(b.detach() if torch.is_tensor(b) else b),
                               rtol=rtol, atol=atol)


@pytest.mark.parametrize("seed", range(6))
def test_linear_train_eval_kl_sweep(ref, seed):
    _, Linear, _ = ref
    rng = np.random.default_rng(seed)
    g = torch.Generator().manual_seed(seed)
    B, K, N, bias = int(rng.integers(1, 40)), int(rng.integers(1, 130)), int(rng.integers(1, 90)), bool(rng.integers(0, 2))
    prior = float(rng.uniform(0.05, 2.0))
    lin = Linear(K, N, bias, sigma_prior=prior)
    _trained_like(lin, g)
    x = torch.randn(B, K, generator=g).requires_grad_(True)
    gout = torch.randn(B, N, generator=g)
    lin.train()
    torch.manual_seed(1000 + seed)
    y = lin(x)
    y.backward(gout)
    torch.manual_seed(1000 + seed)
    eps = torch.empty(B, N).normal_()
    yo, std = O.lrt_linear_fwd(x.detach(), lin.weight, lin.std, lin.bias, eps)
    close(yo, y)
    grads = O.lrt_linear_bwd(x.detach(), lin.weight, lin.std, eps, std, gout)
    close(grads[0], x.grad, 1e-4, 1e-6)
    close(grads[1], lin.weight.grad, 1e-4, 1e-6)
    close(grads[2], lin.std.grad, 1e-4, 1e-7)
    lin.eval()
    torch.manual_seed(2000 + seed)
    with torch.no_grad():
        ye = lin(x)
    torch.manual_seed(2000 + seed)
    close(O.eval_linear_fwd(x.detach(), lin.weight, lin.std, lin.bias, torch.empty(N, K).normal_()), ye)
    close(O.kl_divergence(lin.weight, lin.std, prior), lin.get_kl_divergence(), 1e-5, 1e-4)


@pytest.mark.parametrize("seed", range(8))
def test_conv_train_eval_kl_sweep(ref, seed):
    Conv2d, _, _ = ref
    rng = np.random.default_rng(50 + seed)
    g = torch.Generator().manual_seed(50 + seed)
    B, C, N = int(rng.integers(1, 4)), int(rng.integers(1, 13)), int(rng.integers(1, 17))
    k, stride, dil = int(rng.choice([1, 3, 5])), int(rng.integers(1, 4)), int(rng.choice([1, 1, 2]))
    pad = int(rng.integers(0, k // 2 + 2))
    H, W = int(rng.integers(dil * (k - 1) + 1, 15)), int(rng.integers(dil * (k - 1) + 1, 15))
    conv = Conv2d(C, N, (k, k), stride=stride, padding=pad, dilation=dil, bias=False, sigma_prior=0.05)   # bias: conv.py:32 quirk
    _trained_like(conv, g)
    x = torch.randn(B, C, H, W, generator=g).requires_grad_(True)
    conv.train()
    torch.manual_seed(3000 + seed)
    y = conv(x)
    gout = torch.randn(y.shape, generator=g)
    y.backward(gout)
    torch.manual_seed(3000 + seed)
    eps = torch.empty(y.shape).normal_()
    yo, std = O.lrt_conv_fwd(x.detach(), conv.weight, conv.std, None, eps, stride, pad, dil)
    close(yo, y, 1e-5, 1e-6)
    grads = O.lrt_conv_bwd(x.detach(), conv.weight, conv.std, eps, std, gout, stride, pad, dil)
    close(grads[0], x.grad, 1e-4, 1e-6)
    close(grads[1], conv.weight.grad, 1e-4, 1e-5)
    close(grads[2], conv.std.grad, 1e-4, 1e-7)
    conv.eval()
    torch.manual_seed(4000 + seed)
    with torch.no_grad():
        ye = conv(x)
    torch.manual_seed(4000 + seed)
    close(O.eval_conv_fwd(x.detach(), conv.weight, conv.std, None, torch.empty(conv.weight.shape).normal_(), stride, pad, dil), ye, 1e-5, 1e-6)
    close(O.kl_divergence(conv.weight, conv.std, 0.05), conv.get_kl_divergence(), 1e-5, 1e-3)


@pytest.mark.parametrize("seed", range(4))
def test_dropout_sweep(ref, seed):
    _, _, BernoulliDropout = ref
    rng = np.random.default_rng(80 + seed)
    p = float(rng.choice([0.1, 0.15, 0.2, 0.5]))
    d = BernoulliDropout(p)
    g = torch.Generator().manual_seed(80 + seed)
    for shape in ((int(rng.integers(1, 9)), int(rng.integers(1, 20)), int(rng.integers(1, 7)), int(rng.integers(1, 7))),
                  (int(rng.integers(1, 30)), int(rng.integers(1, 60)))):
        x = torch.randn(*shape, generator=g)
        torch.manual_seed(5000 + seed)
        y = d(x)
        torch.manual_seed(5000 + seed)
        mask = torch.empty(shape[:2] if len(shape) > 2 else shape).bernoulli_(1.0 - d.p)    # tensor-p overload, as dropout.py:21-30
        close(O.dropout_fwd(x, mask, p), y, 1e-6, 0)
