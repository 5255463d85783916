"""CPU (pytest -m "not gpu"): host-side behaviour of the drop-in modules that needs no kernel — construction and seeded
initialisation, the exact-type fused containers, BatchNorm folding, the QAT swap done by prepare_model and the BN
freeze/update API.  The reference behaviour each check follows is cited inline."""
import copy
import pickle

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

import qbn_b200  # noqa: F401
from qbn_b200 import quant_utils as qu, zoo
from qbn_b200.stochastic.bbb import conv as C, linear as L
from qbn_b200.stochastic.bbb.quantized import conv_qat as CQ, linear_qat as LQ
from qbn_b200.stochastic.mcdropout.dropout import BernoulliDropout


def test_initialisation_follows_the_reference_recipe():
    # src/models/stochastic/bbb/conv.py:15-18, linear.py:11-19
    torch.manual_seed(3)
    c = C.Conv2d(3, 8, 3, padding=1, sigma_prior=-2)
    assert c.weight.abs().max() <= 0.01 and torch.all(c.std == -10) and c.bias is None
    assert c.std_prior.dtype == torch.int64 and c.std_prior.item() == -2 and not c.std_prior.requires_grad
    lin = L.Linear(16, 4, True, sigma_prior=0.5)
    assert lin.weight.abs().max() <= 0.01 and torch.all(lin.std == -3) and lin.bias.abs().max() <= 0.01
    assert lin.std_prior.dtype == torch.float32 and lin.std_prior.item() == 0.5
    assert sorted(n for n, _ in lin.named_parameters()) == ["bias", "std", "std_prior", "weight"]
    # seeded construction consumes the generator like the reference: mu ~ U, then rho via uniform_(a, a), then the bias
    # (nn.Conv2d's own reset_parameters runs first, so replay it on a plain conv before comparing)
    torch.manual_seed(3)
    nn.Conv2d(3, 8, 3, padding=1, bias=False)
    w = torch.empty(8, 3, 3, 3).uniform_(-0.01, 0.01)
    torch.empty(8, 3, 3, 3).uniform_(-10, -10)
    after = torch.rand(4)
    torch.manual_seed(3)
    c2 = C.Conv2d(3, 8, 3, padding=1)
    assert torch.equal(c2.weight, w) and torch.equal(torch.rand(4), after)
    assert c._qbn_layer_id != c2._qbn_layer_id


def test_unsupported_conv_options_fail_loudly():
    with pytest.raises(NotImplementedError):
        C.Conv2d(4, 4, 3, groups=2)
    with pytest.raises(NotImplementedError):
        C.Conv2d(4, 4, 3, padding_mode="reflect")


def test_fused_containers_accept_exact_types_only():
    # conv.py:49-68, linear.py:54-59
    c, bn, r = C.Conv2d(3, 8, 3), nn.BatchNorm2d(8), nn.ReLU()
    assert [type(m) for m in C.ConvBnReLU2d(c, bn, r)] == [C.Conv2d, nn.BatchNorm2d, nn.ReLU]
    for bad in (lambda: C.ConvBn2d(nn.Conv2d(3, 8, 3), bn), lambda: C.ConvReLU2d(c, nn.ReLU6()), lambda: C.ConvBnReLU2d(c, bn),
                lambda: L.LinearReLU(nn.Linear(4, 4), r), lambda: C.ConvBn2d(CQ.Conv2d(3, 8, 3, qconfig=qu.qconfig_for(_args())), bn)):
        with pytest.raises(AssertionError, match="Incorrect types"):
            bad()
    box = C.ConvBn2d(c, bn)
    assert isinstance(box, nn.Sequential) and type(box).__name__ == "ConvBn2d" and type(box).__module__ == C.__name__
    clone = pickle.loads(pickle.dumps(box))
    assert type(clone) is C.ConvBn2d and torch.equal(clone[0].weight, c.weight)
    assert type(copy.deepcopy(L.LinearReLU(L.Linear(4, 4, False), r))) is L.LinearReLU


def _args():
    return zoo.Args(sigma_prior=0.1, model="conv_lenet_bbb", q=True, at=True, activation_precision=7, weight_precision=8)


def _random_bn(n):
    bn = nn.BatchNorm2d(n)
    with torch.no_grad():
        bn.running_mean.normal_()
        bn.running_var.uniform_(0.5, 2.0)
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_()
    return bn


@pytest.mark.parametrize("bias", [False, True])
def test_batchnorm_fold_matches_the_closed_form(bias):
    # conv.py:70-88: mu and sigma scale by c = gamma/sqrt(var+eps) per output channel, bias' = (b - mean)*c + beta
    torch.manual_seed(0)
    c, bn = C.Conv2d(3, 8, 3, padding=1, bias=bias).eval(), _random_bn(8).eval()
    with torch.no_grad():
        c.std.uniform_(-4, -1)
    folded = C.fuse_conv_bn(c, bn)
    assert type(folded) is C.Conv2d and folded is not c and folded._qbn_layer_id == c._qbn_layer_id
    k = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).double()
    b0 = c.bias.double() if bias else torch.zeros(8, dtype=torch.double)
    torch.testing.assert_close(folded.weight.double(), c.weight.double() * k.view(-1, 1, 1, 1), rtol=1e-6, atol=1e-9)
    torch.testing.assert_close(F.softplus(folded.std).double(), F.softplus(c.std).double() * k.view(-1, 1, 1, 1), rtol=2e-5, atol=1e-9)
    torch.testing.assert_close(folded.bias.double(), (b0 - bn.running_mean.double()) * k + bn.bias.double(), rtol=1e-5, atol=1e-6)
    with_relu = C.fuse_conv_bn_relu(c, bn, nn.ReLU().eval())
    assert type(with_relu) is C.ConvReLU2d and torch.equal(with_relu[0].weight, folded.weight)


def test_fuse_hooks_by_mode():
    # conv.py:90-115
    c, bn, r = C.Conv2d(3, 8, 3), nn.BatchNorm2d(8), nn.ReLU()
    assert type(C.fuse_conv_bn(c, bn)) is C.ConvBn2d and type(C.fuse_conv_bn_relu(c, bn, r)) is C.ConvBnReLU2d
    with pytest.raises(AssertionError, match="same mode"):
        C.fuse_conv_bn(c, copy.deepcopy(bn).eval())
    with pytest.raises(AssertionError, match="num_features"):
        C.fuse_conv_bn(c, nn.BatchNorm2d(4))
    with pytest.raises(AssertionError, match="affine"):
        C.fuse_conv_bn(c, nn.BatchNorm2d(8, affine=False))
    with pytest.raises(AssertionError, match="eval"):
        C.fuse_conv_bn_eval(c, bn)
    with pytest.raises(NotImplementedError):
        C.fuse_conv_bn_relu(nn.Conv2d(3, 8, 3), bn, r)


def test_prepare_model_swaps_in_the_qat_classes_and_shares_state():
    # quant_utils.py:112-147, linear_qat.py:46-70, conv_qat.py:52-80,169-208
    args = _args()
    c, bn, lin = C.Conv2d(3, 8, 3, padding=1), _random_bn(8), L.Linear(8, 4, False)
    net = nn.Sequential()
    net.add_module("block", C.fuse_conv_bn_relu(c, bn, nn.ReLU()))
    net.add_module("plain", C.Conv2d(8, 8, 1))
    net.add_module("head", L.LinearReLU(lin, nn.ReLU()))
    qu.prepare_model(net.train(), args)
    assert [type(m) for m in net] == [CQ.ConvBnReLU2d, CQ.Conv2d, LQ.LinearReLU]
    assert [m._get_name() for m in net] == ["QATConvBnReLU2d", "QATConv2d", "QATLinearReLU"]
    q = net.block
    assert q.weight is c.weight and q.std is c.std and q.std_prior is c.std_prior and q._qbn_layer_id == c._qbn_layer_id
    assert q.bn.weight is bn.weight and q.bn.running_var is bn.running_var and q.bn.eps == bn.eps
    assert net.head.weight is lin.weight and net.head.add_weight is lin.add_weight
    for m in net:
        assert {"weight_fake_quant", "std_fake_quant", "activation_post_process"} <= set(dict(m.named_children()))
        assert isinstance(m.add_weight.activation_post_process, qu.FakeQuantize)
        assert m.weight_fake_quant.quant_min == -128 and m.activation_post_process.quant_max == 127
    with pytest.raises(AssertionError, match="from_float only works for"):
        CQ.ConvBn2d.from_float(C.Conv2d(3, 8, 3))
    with pytest.raises(AssertionError, match="qconfig"):
        LQ.Linear(4, 4)


def test_qat_batchnorm_freeze_api():
    # conv_qat.py:97-137
    q = CQ.ConvBn2d(3, 8, 3, qconfig=qu.qconfig_for(_args()))
    assert q.freeze_bn is False and q.bn.training
    assert q.freeze_bn_stats() is q and q.freeze_bn and not q.bn.training
    q.train()
    assert q.training and not q.bn.training            # a frozen BN ignores .train()
    assert q.update_bn_stats() is q and q.bn.training
    q.eval()
    assert not q.training and not q.bn.training
    q.train()
    assert q.bn.training
    frozen = CQ.ConvBnReLU2d(3, 8, 3, freeze_bn=True, qconfig=qu.qconfig_for(_args()))
    assert frozen.freeze_bn and not frozen.bn.training
    q.bn.running_mean.fill_(1.0)
    q.reset_running_stats()
    assert torch.all(q.bn.running_mean == 0)


def test_bernoulli_dropout_state_and_identity():
    # src/models/stochastic/mcdropout/dropout.py:9-17
    d = BernoulliDropout(0.25)
    assert set(d.state_dict()) == {"p", "multiplier"} and d.p.item() == 0.25
    torch.testing.assert_close(d.multiplier, torch.tensor([1 / 0.75]))
    x = torch.randn(4, 3)
    assert BernoulliDropout(0.0).eval()(x) is x
    assert "p=0.25" in repr(d)


# ---- int8 checkpoint format: the reference's converted state-dict loads into the drop-in skeleton --------
def _int8_skeleton_cpu(g, args):
    """The converted tiny ResNet with empty int8 modules on the CPU (state-dict plumbing only — no kernel runs here)."""
    from test_gpu_quant_lifecycle import _set_module, _tiny_resnet
    from qbn_b200.stochastic.bbb.quantized import conv_q, linear_q
    net = _tiny_resnet(args).eval()
    net.fuse_model()
    for n in [str(v) for v in g["q_names"]]:
        old = net.get_submodule(n)
        old = old[0] if isinstance(old, nn.Sequential) else old
        relu = bool(g[n + ".relu"])
        if hasattr(old, "in_channels"):
            cls = conv_q.ConvReLU2d if relu else conv_q.Conv2d
            new = cls(old.in_channels, old.out_channels, old.kernel_size, old.stride, old.padding, old.dilation, bias=True, args=args, device="cpu")
        else:
            new = (linear_q.LinearReLU if relu else linear_q.Linear)(old.in_features, old.out_features, args=args, device="cpu")
        _set_module(net, n, new)
    for blk in ("layers.3.0", "layers.3.1"):
        _set_module(net, blk + ".add.add", qu.QFunctional())
    net.quant, net.dequant = qu.Quantize(1.0, 0), qu.DeQuantize()
    return net


def test_reference_int8_checkpoint_loads_into_the_skeleton(golden, golden_dir):
    g = golden("tiny_resnet_int8")
    path = golden_dir / "tiny_resnet_int8_weights.pt"
    ref_sd = torch.load(path, map_location="cpu")
    net = _int8_skeleton_cpu(g, _args())
    assert set(net.state_dict().keys()) == set(ref_sd.keys())          # identical key set, incl. add_weight.* / mul_noise.*
    qu.load_model(net, str(path))
    mods = dict(net.named_modules())
    for n in [str(v) for v in g["q_names"]]:
        m = mods[n]
        assert m.weight.dtype == torch.int8 and np.array_equal(m.weight.numpy(), g[n + ".mu_q"])
        assert np.array_equal(m.std.numpy(), g[n + ".sigma_q"])
        for attr in ("mu_qp", "sigma_qp", "mul_qp", "add_qp"):
            want = g["%s.%s" % (n, attr)]
            assert getattr(m, attr) == (float(want[0]), int(want[1])), (n, attr)
        assert (m.scale, m.zero_point) == (float(g[n + ".out_qp"][0]), int(g[n + ".out_qp"][1]))
        bias = g[n + ".bias"]
        assert (m.bias() is None) if bias.size == 0 else np.array_equal(m.bias().numpy(), bias)
    for blk in ("layers.3.0", "layers.3.1"):
        f = mods[blk + ".add.add"]
        assert (f.scale, f.zero_point) == (float(g[blk + ".add.y_qp"][0]), int(g[blk + ".add.y_qp"][1]))
    assert (net.quant.scale, net.quant.zero_point) == (float(g["quant_qp"][0]), int(g["quant_qp"][1]))
    # and back: what the drop-in saves is what the reference saved (values and dtypes)
    out = net.state_dict()
    for k, v in ref_sd.items():
        if v is None:
            assert out[k] is None
        elif v.is_quantized:
            assert out[k].is_quantized and torch.equal(out[k].int_repr(), v.int_repr()), k
            assert out[k].q_scale() == v.q_scale() and out[k].q_zero_point() == v.q_zero_point(), k
        else:
            assert out[k].dtype == v.dtype and out[k].shape == v.shape and torch.equal(out[k].detach(), v.detach()), k


def test_sample_batch_context_and_engine_guards():
    from qbn_b200 import noise
    from qbn_b200.mc_int8 import Int8MCEngine, balanced_chunks
    assert noise.sample_batch_state() is None
    with noise.sample_batch(5, 40, 256, act_bits=7):
        assert noise.sample_batch_state() == (5, 40, 256, 7)
        with noise.sample_batch(2, 0, 8):
            assert noise.sample_batch_state() == (2, 0, 8, 8)
        assert noise.sample_batch_state() == (5, 40, 256, 7)
    assert noise.sample_batch_state() is None
    with noise.inject([torch.zeros(1)]):
        with pytest.raises(RuntimeError, match="injected noise"):
            with noise.sample_batch(2, 0, 8):
                pass
    # chunks: sizes differ by at most one, never exceed the limit, cover every sample once
    for total, chunk in ((100, 25), (100, 30), (13, 50), (7, 1), (12, 5)):
        parts = balanced_chunks(total, chunk)
        assert sum(parts) == total and max(parts) <= chunk and max(parts) - min(parts) <= 1
    assert balanced_chunks(0, 4) == []
    with pytest.raises(ValueError, match="converted model"):
        Int8MCEngine(nn.Sequential(nn.Linear(4, 4)))
    from qbn_b200.stochastic.bbb.quantized import linear_q
    q = nn.Sequential(linear_q.Linear(4, 4, device="cpu"), BernoulliDropout(0.5))
    assert Int8MCEngine(q).act_bits == 8                           # int8 MC-Dropout sites are driven by the engine (dropout.py:31-39)
    eng = Int8MCEngine(nn.Sequential(linear_q.Linear(4, 4, device="cpu")))
    assert eng.act_bits == 8 and not eng.regression                # no args on the model: nothing is assumed about its clamping
    with pytest.raises(RuntimeError, match="CUDA"):
        eng.predict_sum(torch.zeros(2, 4), 3)
    # the layout rule of the sample-batched layers
    from qbn_b200.quant_utils import QTensor
    lay = linear_q.Linear._batch_layout
    assert lay(QTensor(torch.zeros(8, 4, dtype=torch.uint8), 1.0, 0), 3, 8) is True
    assert lay(QTensor(torch.zeros(24, 4, dtype=torch.uint8), 1.0, 0), 3, 8) is False
    with pytest.raises(ValueError, match="leading dimension"):
        lay(QTensor(torch.zeros(16, 4, dtype=torch.uint8), 1.0, 0), 3, 8)


def test_losses_match_the_reference(golden):
    """src/losses.py through the mirror: ELBO value, its two terms and the gradient w.r.t. the model output, for both tasks
    and both scalings (fixtures from the reference, oracle/make_golden.py:gen_losses)."""
    from qbn_b200 import losses as L
    g = golden("losses")
    args = zoo.Args(loss_multiplier=0.5)
    out, tgt = torch.as_tensor(g["out"]), torch.as_tensor(g["target"])
    loss, ce, kl = L.ClassificationLoss(zoo.Args(loss_multiplier=1.0), "batch")(out, tgt, torch.tensor(123.4), 0.01, 176, 45000)
    for got, key in ((loss, "loss"), (ce, "ce"), (kl, "kl")):
        torch.testing.assert_close(got, torch.as_tensor(g[key]), rtol=1e-6, atol=1e-7)
    mean, var, rtgt = (torch.as_tensor(g[k]) for k in ("r_mean", "r_var", "r_target"))
    for scaling in ("batch", "whole"):
        o = out.clone().requires_grad_(True)
        vals = L.LOSS_FACTORY["classification"](args, scaling)(o, tgt, torch.tensor(123.4), 0.01, 176, 45000)
        vals[0].backward()
        np.testing.assert_allclose([v.item() for v in vals], g["cls_%s" % scaling], rtol=1e-6)
        np.testing.assert_allclose(o.grad.numpy(), g["cls_%s_dout" % scaling], rtol=1e-5, atol=1e-7)
        m, v = mean.clone().requires_grad_(True), var.clone().requires_grad_(True)
        vals = L.LOSS_FACTORY["regression"](args, scaling)((m, v), rtgt, torch.tensor(55.5), 0.1, 8, 1000)
        vals[0].backward()
        np.testing.assert_allclose([v.item() for v in vals], g["reg_%s" % scaling], rtol=1e-5)
        np.testing.assert_allclose(m.grad.numpy(), g["reg_%s_dmean" % scaling], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(v.grad.numpy(), g["reg_%s_dvar" % scaling], rtol=1e-4, atol=1e-6)
    with pytest.raises(NotImplementedError):
        L.ClassificationLoss(args, "per-sample")(out, tgt, torch.tensor(1.0), 0.01, 176, 45000)
