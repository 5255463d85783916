"""CPU (pytest -m "not gpu"): the host-side logic of the int8 path with the kernels replaced by oracle-backed stand-ins
(tests/_i8_emulation.py) — module plumbing, QTensor flow through the model glue, checkpoint loading, sample batching.
The GPU suite proves the kernels; this proves, without a GPU, that the Python around them composes a network the way the
reference does (every intermediate integer map of the reference's FBGEMM forward is reproduced)."""
import numpy as np
import pytest
import torch
import torch.nn as nn

import qbn_b200  # noqa: F401
from qbn_b200 import noise, quant_utils as qu, zoo

from _i8_emulation import emulated_int8_ops
from test_modules_cpu import _args, _int8_skeleton_cpu


@pytest.fixture
def net(golden, golden_dir, monkeypatch):
    emulated_int8_ops(monkeypatch)
    g = golden("tiny_resnet_int8")
    m = _int8_skeleton_cpu(g, _args())
    qu.load_model(m, str(golden_dir / "tiny_resnet_int8_weights.pt"))
    return m.eval(), g


def test_whole_int8_network_through_the_modules(net):
    m, g = net
    q_names = [str(n) for n in g["q_names"]]
    mods, seen = dict(m.named_modules()), {}
    watch = [str(n) for n in g["order"] if str(n) != "layers.4"]
    hooks = [mods[n].register_forward_hook(lambda mod, i, o, n=n: seen.__setitem__(n, (i, o))) for n in watch]
    with torch.no_grad(), noise.inject([torch.as_tensor(g[n + ".eps"]) for n in q_names]):
        y = m(torch.as_tensor(g["x"]))
    for h in hooks:
        h.remove()
    for n in watch:
        out = seen[n][1]
        assert np.array_equal(out.q.numpy().reshape(g[n + ".y_q"].shape), g[n + ".y_q"]), n
        assert out.zero_point == int(g[n + ".y_qp"][1]) and abs(out.scale - g[n + ".y_qp"][0]) < 1e-12, n
    assert np.array_equal(seen["layers.6"][0][0].q.numpy(), g["layers.6.x_q"])
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("bits", [8, 7])
def test_sample_batched_forward_equals_the_per_sample_loop(net, bits):
    m, g = net
    x = torch.as_tensor(g["x"])
    noise.manual_seed(123)
    S, B = 3, x.shape[0]
    with torch.no_grad():
        loop = []
        for s in range(S):
            with noise.sample_index(5 + s):
                loop.append(m(x))
        with noise.sample_batch(S, 5, B, act_bits=bits):
            batched = m(x)
    assert batched.shape == (S * B, 10)
    assert float((loop[0] - loop[1]).abs().max()) > 0
    for s in range(S):
        assert torch.equal(batched[s * B:(s + 1) * B], loop[s]), s          # same integers -> same probabilities


def test_engine_mode_skips_redundant_activation_clamps(net, monkeypatch):
    """clamp_activation runs after every module (models_bbb.py:172-182); when the producing kernel already clamped to the
    model's activation width the pass is the identity and is skipped: in engine mode only the quantised input needs one."""
    m, g = net
    calls = []
    real = torch.clamp
    monkeypatch.setattr(torch, "clamp", lambda *a, **k: (calls.append(1), real(*a, **k))[1])
    from qbn_b200 import ops
    relus, adds = [], []
    real_relu, real_add = ops.i8_relu, ops.i8_add
    monkeypatch.setattr(ops, "i8_relu", lambda *a, **k: (relus.append(1), real_relu(*a, **k))[1])
    monkeypatch.setattr(ops, "i8_add", lambda *a, **k: (adds.append(k.get("relu", False)), real_add(*a, **k))[1])
    x = torch.as_tensor(g["x"])
    noise.manual_seed(9)
    with torch.no_grad(), noise.sample_batch(2, 0, x.shape[0], act_bits=7):
        m(x)
    in_engine = (len(calls), len(relus), list(adds))
    calls.clear(), relus.clear(), adds.clear()
    with torch.no_grad(), noise.sample_index(0):
        m(x)
    per_sample = (len(calls), len(relus), list(adds))
    # engine: one clamp (the input, quantised to the full uint8 range); each block's add + ReLU + clamp is ONE launch
    assert in_engine == (1, 0, [True, True]), in_engine
    # module mode keeps the reference's sequence: + a clamp per int8 layer (8-bit outputs) and per add, separate ReLU launches
    assert per_sample == (1 + 7 + 2, 2, [False, False]), per_sample


# ---- LeNet-shaped net (max-pooling between the convolutions, Linear+ReLU fused): tests/golden/tiny_int8.npz ---------------
def _lenet_int8(g, args, device="cpu"):
    from test_gpu_quant_lifecycle import _set_module, _tiny_net
    from qbn_b200.stochastic.bbb.quantized import conv_q, linear_q
    m = _tiny_net(args).eval()
    m.fuse_model()
    for n in [str(v) for v in g["q_names"]]:
        mu_q, relu = g[n + ".mu_q"], bool(g[n + ".relu"])
        if mu_q.ndim == 4:
            stride, pad = [int(v) for v in g[n + ".conv"]]
            new = (conv_q.ConvReLU2d if relu else conv_q.Conv2d)(mu_q.shape[1], mu_q.shape[0], mu_q.shape[2:], stride=(stride, stride),
                                                                  padding=(pad, pad), dilation=(1, 1), args=args, device=device)
        else:
            new = (linear_q.LinearReLU if relu else linear_q.Linear)(mu_q.shape[1], mu_q.shape[0], args=args, device=device)
        new.weight, new.std = torch.as_tensor(mu_q).to(device), torch.as_tensor(g[n + ".sigma_q"]).to(device)
        for attr in ("mu_qp", "sigma_qp", "mul_qp", "add_qp"):
            setattr(new, attr, (float(g["%s.%s" % (n, attr)][0]), int(g["%s.%s" % (n, attr)][1])))
        new.scale, new.zero_point = float(g[n + ".out_qp"][0]), int(g[n + ".out_qp"][1])
        _set_module(m, n, new)
    m.quant, m.dequant = qu.Quantize(float(g["quant_qp"][0]), int(g["quant_qp"][1])), qu.DeQuantize()
    return m


def test_lenet_shaped_int8_network_and_its_sample_batching(golden, monkeypatch):
    emulated_int8_ops(monkeypatch)
    g = golden("tiny_int8")
    m = _lenet_int8(g, _args())
    q_names = [str(n) for n in g["q_names"]]
    x = torch.as_tensor(g["x"])
    with torch.no_grad(), noise.inject([torch.as_tensor(g[n + ".eps"]) for n in q_names]):
        y = m(x)
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=1e-5, atol=1e-7)       # the reference's int8 forward, end to end
    noise.manual_seed(4)
    S, B = 3, x.shape[0]
    with torch.no_grad():
        loop = []
        for s in range(S):
            with noise.sample_index(s):
                loop.append(m(x))
        with noise.sample_batch(S, 0, B, act_bits=7):
            batched = m(x)
    for s in range(S):
        assert torch.equal(batched[s * B:(s + 1) * B], loop[s]), s
