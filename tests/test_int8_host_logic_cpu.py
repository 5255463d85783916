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
