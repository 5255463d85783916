"""CPU (pytest -m "not gpu"): the reference's OWN stock-quantised model code (MC-Dropout ResNet, SGHMC ensemble: models_mc.py,
models_sgld.py, src/utils.py:25-55, dropout.py) running on the drop-in int8 modules after quant_utils.to_device_int8 — with
the kernels replaced by the oracle-backed stand-ins of tests/_i8_emulation.py — against the same reference model on torch's
FBGEMM kernels.  What this pins without a GPU: the module transplant (weights, qparams, biases), QTensor standing in for a
torch quint8 tensor inside the reference's forward (clamp_activation, ReLU, AvgPool2d, Flatten, Add), the int8 MC-Dropout
branch at the module level, and the ensemble's member cycling.  Skipped where the reference is absent."""
import numpy as np
import pytest
import torch

import qbn_b200  # noqa: F401
from qbn_b200 import noise, quant_utils as qu

import _ref_models as R
from _i8_emulation import emulated_int8_ops

pytestmark = pytest.mark.skipif(not R.available(), reason="reference sources not present (neither /root/reference nor oracle/_ref)")


def test_reference_sghmc_ensemble_on_dropin_modules(monkeypatch):
    emulated_int8_ops(monkeypatch)
    torch.set_num_threads(1)
    net, args = R.sgld_ensemble(2)
    x = torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(3))
    mine = qu.to_device_int8(R.clone(net), "cpu")
    with torch.no_grad():
        want = [net(x) for _ in range(2)]                 # Network.forward cycles the members (models_sgld.py:277-284)
        got = [mine(x) for _ in range(2)]
    assert float((want[0] - want[1]).abs().max()) > 0      # the members really differ
    for w, g in zip(want, got):
        np.testing.assert_allclose(g.numpy(), w.numpy(), rtol=1e-6, atol=1e-7)
    # every int8 module of the transplant holds the reference's integers
    from qbn_b200.stochastic.quantized_det import QuantizedConv2d
    ref_mods, my_mods = dict(net.named_modules()), dict(mine.named_modules())
    n_conv = 0
    for name, m in my_mods.items():
        if isinstance(m, QuantizedConv2d):
            r = ref_mods[name]
            assert torch.equal(m.weight, r.weight().int_repr()) and m.w_qp == (float(r.weight().q_scale()), int(r.weight().q_zero_point()))
            assert (m.scale, m.zero_point) == (float(r.scale), int(r.zero_point))
            n_conv += 1
    assert n_conv == 2 * 20


def test_reference_mc_dropout_resnet_on_dropin_modules(monkeypatch):
    emulated_int8_ops(monkeypatch)
    torch.set_num_threads(1)
    net, args = R.mc_dropout_resnet()
    x = torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(4))
    sites = R.dropout_sites(net)
    assert len(sites) == 20                                # models_mc.py:129-140,180
    shapes = []
    hooks = [m.register_forward_hook(lambda mod, i, o: shapes.append(tuple(i[0].shape[:2]))) for m in sites]
    torch.manual_seed(77)
    with torch.no_grad():
        want = net(x)
    for h in hooks:
        h.remove()
    torch.manual_seed(77)                                  # replay: one mask per site in forward order (dropout.py:19-30)
    masks = [torch.FloatTensor(*s).bernoulli_(1. - sites[0].p) for s in shapes]      # tensor-p overload, exactly like dropout.py:21
    mine = qu.to_device_int8(R.clone(net), "cpu")
    with torch.no_grad(), noise.inject(masks):
        got = mine(x)
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-6, atol=1e-7)
