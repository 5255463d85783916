"""GPU (pytest -m gpu): the quantisation lifecycle of config C3 on the drop-in modules — prepare_model
(QAT swap + observers), one train forward, one eval forward (calibrates add_weight/mul_noise), convert,
int8 forward — against the reference run captured in tests/golden/tiny_{qat,int8}.npz.

Fake-quantised values sit on a grid; an fp32 re-association upstream can move a pre-quantisation value
across a rounding boundary, so float comparisons allow a handful of one-quantum differences."""
import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def _tiny_net(args):
    from qbn_b200 import zoo
    from qbn_b200.stochastic.bbb.conv import Conv2d
    from qbn_b200.stochastic.bbb.linear import Linear
    net = zoo.ConvNetwork_LeNet([1, 1, 28, 28], 10, True, args)
    sp = args.sigma_prior
    net.layers = nn.ModuleList([
        Conv2d(1, 6, (5, 5), stride=1, padding=2, sigma_prior=sp, bias=False, args=args), nn.MaxPool2d(2, 2),
        Conv2d(6, 12, (5, 5), stride=1, padding=2, sigma_prior=sp, bias=False, args=args), nn.MaxPool2d(2, 2),
        zoo.Flatten(), Linear(12 * 7 * 7, 32, sigma_prior=sp, bias=False, args=args), nn.ReLU(),
        Linear(32, 10, sigma_prior=sp, bias=False, args=args)])
    return net


def mostly_equal(got, ref, quantum, frac=0.995):
    got = got.detach().float().cpu().numpy()
    ref = np.asarray(ref, dtype=np.float32)
    diff = np.abs(got - ref)
    ok = diff <= 1e-5 * max(1.0, float(np.abs(ref).max()))
    assert ok.mean() >= frac, "only %.4f of the elements agree" % ok.mean()
    assert diff.max() <= 2.5 * quantum + 1e-6, "max diff %g vs quantum %g" % (diff.max(), quantum)


def test_qat_then_int8_lifecycle(golden):
    import __graft_entry__ as ge
    ge.build()
    from qbn_b200 import noise, quant_utils as qu, zoo
    g = golden("tiny_qat")
    args = zoo.Args(sigma_prior=0.1, model="conv_lenet_bbb", q=True, at=True, activation_precision=7, weight_precision=8)
    net = _tiny_net(args)
    names = [str(n) for n in g["qat_names"]]
    mods = dict(net.named_modules())
    with torch.no_grad():
        for n in names:
            mods[n].weight.copy_(torch.as_tensor(g["p.%s.weight" % n]))
            mods[n].std.copy_(torch.as_tensor(g["p.%s.std" % n]))
    net.train()
    qu.prepare_model(net, args)
    net = net.cuda()
    mods = dict(net.named_modules())
    assert [type(mods[n]).__name__ for n in names] == ["Conv2d", "Conv2d", "LinearReLU", "Linear"]
    x = torch.as_tensor(g["x"]).cuda()
    caps = {}
    hooks = [mods[n].register_forward_hook(lambda m, i, o, n=n: caps.__setitem__(n, (i[0].detach(), o.detach()))) for n in names]
    # ---- train-mode forward: LRT with fake-quantised (mu~, sigma~), first observer call initialises min/max
    with noise.inject([torch.as_tensor(g["tr.%s.eps" % n]).cuda() for n in names]):
        y = net(x)
    for n in names:
        m = mods[n]
        for key, fq in (("wfq", m.weight_fake_quant), ("sfq", m.std_fake_quant), ("afq", m.activation_post_process)):
            s_ref, z_ref = g["tr.%s.%s" % (n, key)][:2]
            np.testing.assert_allclose(float(fq.scale), s_ref, rtol=2e-5)
            assert abs(int(fq.zero_point) - int(z_ref)) <= 1
        mostly_equal(caps[n][1], g["tr.%s.out" % n], float(m.activation_post_process.scale))
    mostly_equal(y, g["tr.y"], 1e-3)
    y.sum().backward()                                              # STE + LRT backward run end to end
    assert all(torch.isfinite(mods[n].weight.grad).all() and torch.isfinite(mods[n].std.grad).all() for n in names)
    # ---- eval-mode forward: EMA observer update, calibrates the add_weight / mul_noise fake-quants
    net.eval()
    with torch.no_grad(), noise.inject([torch.as_tensor(g["ev.%s.eps" % n]).cuda() for n in names]):
        ye = net(x)
    for n in names:
        m = mods[n]
        for key, fq in (("mulfq", m.mul_noise.activation_post_process), ("addfq", m.add_weight.activation_post_process),
                        ("afq", m.activation_post_process)):
            s_ref, z_ref = g["ev.%s.%s" % (n, key)][:2]
            np.testing.assert_allclose(float(fq.scale), s_ref, rtol=5e-4)
            assert abs(int(fq.zero_point) - int(z_ref)) <= 1
        mostly_equal(caps[n][1], g["ev.%s.out" % n], float(m.activation_post_process.scale), frac=0.98)
    mostly_equal(ye, g["ev.y"], 2e-3, frac=0.98)
    for h in hooks:
        h.remove()
    # ---- convert to int8 and run with replayed noise
    gi = golden("tiny_int8")
    qu.convert(net)
    net.eval()
    mods = dict(net.named_modules())
    assert [type(mods[n]).__name__ for n in names] == ["Conv2d", "Conv2d", "LinearReLU", "Linear"]
    assert type(net.quant).__name__ == "Quantize"
    np.testing.assert_allclose(net.quant.scale, gi["quant_qp"][0], rtol=1e-5)
    for n in names:
        m = mods[n]
        np.testing.assert_allclose(m.mu_qp[0], gi[n + ".mu_qp"][0], rtol=5e-4)
        np.testing.assert_allclose(m.sigma_qp[0], gi[n + ".sigma_qp"][0], rtol=5e-4)
        np.testing.assert_allclose(m.scale, gi[n + ".out_qp"][0], rtol=5e-4)
        agree = (m.weight.cpu().numpy() == gi[n + ".mu_q"]).mean()
        assert agree > 0.97, (n, agree)
    with torch.no_grad(), noise.inject([torch.as_tensor(gi[n + ".eps"]).cuda() for n in names]):
        yq = net(x)
    assert yq.shape == (8, 10) and torch.isfinite(yq).all()
    assert float((yq.cpu() - torch.as_tensor(gi["y"])).abs().max()) < 0.05           # same predictive distribution
    # state-dict round trip with the reference's key set (conv_q.py:72-78)
    sd = mods[names[0]].state_dict()
    assert {"scale", "zero_point", "weight", "std", "bias_"} <= set(sd.keys())
    assert sd["weight"].dtype == torch.qint8


def test_int8_modules_bit_exact_with_reference_qparams(golden):
    """Drop-in int8 modules loaded with the reference's own int8 state (mu_q, sigma_q, every qparam):
    sampled weights and layer outputs are bit-identical to the reference's FBGEMM forward."""
    import __graft_entry__ as ge
    ge.build()
    from qbn_b200 import noise, zoo
    from qbn_b200.quant_utils import QTensor
    from qbn_b200.stochastic.bbb.quantized import conv_q, linear_q
    g = golden("tiny_int8")
    args = zoo.Args(activation_precision=7, weight_precision=8)
    for n in [str(v) for v in g["q_names"]]:
        mu_q, relu = g[n + ".mu_q"], bool(g[n + ".relu"])
        if mu_q.ndim == 4:
            stride, pad = [int(v) for v in g[n + ".conv"]]
            cls = conv_q.ConvReLU2d if relu else conv_q.Conv2d
            m = cls(mu_q.shape[1], mu_q.shape[0], mu_q.shape[2:], stride=(stride, stride), padding=(pad, pad), dilation=(1, 1), args=args)
        else:
            cls = linear_q.LinearReLU if relu else linear_q.Linear
            m = cls(mu_q.shape[1], mu_q.shape[0], args=args)
        m.weight, m.std = torch.as_tensor(mu_q).cuda(), torch.as_tensor(g[n + ".sigma_q"]).cuda()
        m.mu_qp = (float(g[n + ".mu_qp"][0]), int(g[n + ".mu_qp"][1]))
        m.sigma_qp = (float(g[n + ".sigma_qp"][0]), int(g[n + ".sigma_qp"][1]))
        m.mul_qp = (float(g[n + ".mul_qp"][0]), int(g[n + ".mul_qp"][1]))
        m.add_qp = (float(g[n + ".add_qp"][0]), int(g[n + ".add_qp"][1]))
        m.scale, m.zero_point = float(g[n + ".out_qp"][0]), int(g[n + ".out_qp"][1])
        xq = torch.as_tensor(g[n + ".x_q"]).cuda()
        if xq.dim() == 4:
            xq = xq.contiguous(memory_format=torch.channels_last)
        x = QTensor(xq, float(g[n + ".x_qp"][0]), int(g[n + ".x_qp"][1]))
        with noise.inject([torch.as_tensor(g[n + ".eps"]).cuda()]):
            y = m(x)
        assert np.array_equal(y.q.cpu().numpy(), g[n + ".y_q"]), n
        assert abs(y.scale - g[n + ".y_qp"][0]) < 1e-12 and y.zero_point == int(g[n + ".y_qp"][1])


# ---- ResNet-shaped net: conv+BN(+ReLU) fusion, BN fold at convert, quantised residual add / ReLU / average pool ----------
def _tiny_resnet(args):
    from qbn_b200 import zoo
    from qbn_b200.stochastic.bbb.conv import Conv2d
    from qbn_b200.stochastic.bbb.linear import Linear
    net = zoo.ConvNetwork_ResNet([1, 3, 32, 32], 10, True, args)
    sp = args.sigma_prior
    net.layers = nn.ModuleList([
        Conv2d(3, 8, kernel_size=3, stride=1, padding=1, bias=False, sigma_prior=sp, args=args), nn.BatchNorm2d(8), nn.ReLU(),
        nn.ModuleList([zoo.BasicBlock(8, 8, 1, True, args), zoo.BasicBlock(8, 16, 2, True, args)]),
        nn.AvgPool2d(4), zoo.Flatten(), Linear(16, 10, sigma_prior=sp, bias=False, args=args)])
    return net


def _load_float_state(net, g):
    mods = dict(net.named_modules())
    with torch.no_grad():
        for key in g.files:
            if key.startswith("p."):
                name, _, leaf = key[2:].rpartition(".")
                getattr(mods[name], leaf).copy_(torch.as_tensor(g[key]))


def _set_module(net, name, new):
    parent, _, leaf = name.rpartition(".")
    (net.get_submodule(parent) if parent else net)._modules[leaf] = new


def _int8_module(g, n, args):
    """A drop-in int8 layer carrying the reference's own int8 state for node `n` of a fixture."""
    from qbn_b200.stochastic.bbb.quantized import conv_q, linear_q
    mu_q, relu = g[n + ".mu_q"], bool(g[n + ".relu"])
    if mu_q.ndim == 4:
        stride, pad = [int(v) for v in g[n + ".conv"]]
        cls = conv_q.ConvReLU2d if relu else conv_q.Conv2d
        m = cls(mu_q.shape[1], mu_q.shape[0], mu_q.shape[2:], stride=(stride, stride), padding=(pad, pad), dilation=(1, 1), args=args)
    else:
        m = (linear_q.LinearReLU if relu else linear_q.Linear)(mu_q.shape[1], mu_q.shape[0], args=args)
    m.weight, m.std = torch.as_tensor(mu_q).cuda(), torch.as_tensor(g[n + ".sigma_q"]).cuda()
    for attr in ("mu_qp", "sigma_qp", "mul_qp", "add_qp"):
        setattr(m, attr, (float(g["%s.%s" % (n, attr)][0]), int(g["%s.%s" % (n, attr)][1])))
    m.scale, m.zero_point = float(g[n + ".out_qp"][0]), int(g[n + ".out_qp"][1])
    if (n + ".bias") in g.files and g[n + ".bias"].size:          # the folded BatchNorm shift (conv_q.py:130-133)
        m.bias_ = torch.as_tensor(g[n + ".bias"]).cuda()
    return m


def test_resnet_qat_then_int8_lifecycle(golden):
    import __graft_entry__ as ge
    ge.build()
    from qbn_b200 import noise, quant_utils as qu, zoo
    g = golden("tiny_resnet_int8")
    args = zoo.Args(sigma_prior=0.1, model="conv_resnet_bbb", q=True, at=True, activation_precision=7, weight_precision=8)
    net = _tiny_resnet(args)
    _load_float_state(net, g)
    net.train()
    qu.prepare_model(net, args)
    mods = dict(net.named_modules())
    names = [str(n) for n in g["qat_names"]]
    assert [type(mods[n]).__name__ for n in names] == [str(t) for t in g["qat_types"]]
    for m in net.modules():
        if hasattr(m, "freeze_bn_stats"):
            m.freeze_bn_stats()
    net = net.cuda()
    x = torch.as_tensor(g["x"]).cuda()
    with noise.inject([torch.as_tensor(g["tr.%s.eps" % n]).cuda() for n in names]):
        y = net(x)                                       # QAT train forward: LRT over fake-quantised, BN-scaled (mu~, sigma~)
    assert y.shape == (8, 10) and torch.isfinite(y).all()
    assert float((y.detach().cpu() - torch.as_tensor(g["tr.y"])).abs().max()) < 0.05
    y.sum().backward()
    assert all(mods[n].weight.grad is not None and torch.isfinite(mods[n].weight.grad).all() for n in names)
    net.eval()
    with torch.no_grad(), noise.inject([torch.as_tensor(g["ev.%s.eps" % n]).cuda() for n in names]):
        ye = net(x)                                      # calibrates add_weight / mul_noise
    assert float((ye.cpu() - torch.as_tensor(g["ev.y"])).abs().max()) < 0.05
    qu.convert(net)
    net.eval()
    mods = dict(net.named_modules())
    q_names = [str(n) for n in g["q_names"]]
    assert [type(mods[n]).__name__ for n in q_names] == ["ConvReLU2d", "ConvReLU2d", "Conv2d", "ConvReLU2d", "Conv2d", "Conv2d", "Linear"]
    assert type(net.layers[3][0].add.add).__name__ == "QFunctional" and type(net.quant).__name__ == "Quantize"
    np.testing.assert_allclose(net.quant.scale, g["quant_qp"][0], rtol=1e-5)
    for n in q_names:                                    # BN folded into (mu, sigma) before quantisation, like the reference
        np.testing.assert_allclose(mods[n].mu_qp[0], g[n + ".mu_qp"][0], rtol=5e-4)
        np.testing.assert_allclose(mods[n].sigma_qp[0], g[n + ".sigma_qp"][0], rtol=5e-4)
        assert (mods[n].weight.cpu().numpy() == g[n + ".mu_q"]).mean() > 0.97, n
    with torch.no_grad(), noise.inject([torch.as_tensor(g[n + ".eps"]).cuda() for n in q_names]):
        yq = net(x)
    assert yq.shape == (8, 10) and torch.isfinite(yq).all()
    assert float((yq.cpu() - torch.as_tensor(g["y"])).abs().max()) < 0.1


def test_resnet_int8_network_bit_exact_with_reference_state(golden):
    """The whole converted network — quantise, stem conv, identity block, stride-2 block with its 1x1 shortcut, quantised
    ReLU / residual add / average pool, linear, dequantise — loaded with the reference's int8 state and replayed noise:
    every intermediate integer map equals the reference's FBGEMM forward."""
    import __graft_entry__ as ge
    ge.build()
    from qbn_b200 import noise, quant_utils as qu, zoo
    g = golden("tiny_resnet_int8")
    args = zoo.Args(sigma_prior=0.1, model="conv_resnet_bbb", q=True, at=True, activation_precision=7, weight_precision=8)
    net = _tiny_resnet(args).eval()
    net.fuse_model()                                     # eval-mode fusion: BN slots -> Identity
    q_names = [str(n) for n in g["q_names"]]
    for n in q_names:
        _set_module(net, n, _int8_module(g, n, args))
    for blk in ("layers.3.0", "layers.3.1"):
        s, z = g[blk + ".add.y_qp"]
        _set_module(net, blk + ".add.add", qu.QFunctional(float(s), int(z)))
    net.quant, net.dequant = qu.Quantize(float(g["quant_qp"][0]), int(g["quant_qp"][1])), qu.DeQuantize()
    _check_int8_forward_bit_exact(net.cuda(), g)


def _check_int8_forward_bit_exact(net, g):
    """Replay the fixture's noise through the converted net and compare every captured integer map with the reference's."""
    from qbn_b200 import noise
    q_names = [str(n) for n in g["q_names"]]
    mods, seen = dict(net.named_modules()), {}
    watch = [str(n) for n in g["order"] if str(n) != "layers.4"]
    hooks = [mods[n].register_forward_hook(lambda m, i, o, n=n: seen.__setitem__(n, (i, o))) for n in watch]
    with torch.no_grad(), noise.inject([torch.as_tensor(g[n + ".eps"]).cuda() for n in q_names]):
        y = net(torch.as_tensor(g["x"]).cuda())
    for h in hooks:
        h.remove()
    for n in watch:
        out = seen[n][1]
        assert np.array_equal(out.q.cpu().numpy().reshape(g[n + ".y_q"].shape), g[n + ".y_q"]), n
        assert abs(out.scale - g[n + ".y_qp"][0]) < 1e-12 and out.zero_point == int(g[n + ".y_qp"][1]), n
    pooled = seen["layers.6"][0][0]                       # AvgPool2d(4) + Flatten feed the classifier
    assert np.array_equal(pooled.q.cpu().numpy(), g["layers.6.x_q"])
    assert np.array_equal(pooled.q.cpu().numpy().reshape(8, 16, 1, 1), g["layers.4.y_q"])
    np.testing.assert_allclose(y.cpu().numpy(), g["y"], rtol=1e-5, atol=1e-7)


def test_reference_int8_checkpoint_runs_bit_exact(golden, golden_dir):
    """the checkpoint the reference writes after convert (`weights.pt`: qint8 tensors + qparams) loaded the
    way the reference's evaluation scripts do — fresh model -> prepare_model -> convert -> load_model
    (experiments/utils.py:153-158) — runs on the CUDA int8 path and reproduces the reference's integers."""
    import __graft_entry__ as ge
    ge.build()
    from qbn_b200 import quant_utils as qu, zoo
    g = golden("tiny_resnet_int8")
    args = zoo.Args(sigma_prior=0.1, model="conv_resnet_bbb", q=True, at=True, activation_precision=7, weight_precision=8)
    net = _tiny_resnet(args)
    qu.prepare_model(net, args)
    qu.convert(net.cuda())                                # un-calibrated skeleton; every number comes from the file
    qu.load_model(net, str(golden_dir / "tiny_resnet_int8_weights.pt"))
    net.eval()
    assert set(net.state_dict().keys()) == set(torch.load(golden_dir / "tiny_resnet_int8_weights.pt", map_location="cpu").keys())
    _check_int8_forward_bit_exact(net, g)


def _converted_tiny_resnet(golden, golden_dir):
    from qbn_b200 import quant_utils as qu, zoo
    args = zoo.Args(sigma_prior=0.1, model="conv_resnet_bbb", q=True, at=True, activation_precision=7, weight_precision=8)
    net = _tiny_resnet(args)
    qu.prepare_model(net, args)
    qu.convert(net.cuda())
    qu.load_model(net, str(golden_dir / "tiny_resnet_int8_weights.pt"))
    return net.eval()


@pytest.mark.parametrize("tensor_cores", [False, True])
def test_int8_sample_batched_engine_equals_the_per_sample_loop(golden, golden_dir, tensor_cores):
    """All MC samples of a chunk in one forward (Int8MCEngine) == the reference's loop of single forwards with the same
    per-sample Philox streams, bit for bit; independent of the chunk size and of the first sample index (sharding)."""
    import __graft_entry__ as ge
    ge.build()
    from qbn_b200 import noise
    from qbn_b200.mc_int8 import Int8MCEngine
    g = golden("tiny_resnet_int8")
    net = _converted_tiny_resnet(golden, golden_dir)
    x = torch.as_tensor(g["x"]).cuda()
    noise.manual_seed(77)
    S = 6
    with torch.no_grad():
        loop = []
        for s in range(S):
            with noise.sample_index(s):
                loop.append(net(x))
    want = torch.stack(loop).sum(0)
    assert float((loop[0] - loop[1]).abs().max()) > 0            # the samples really differ
    got = Int8MCEngine(net, chunk=4, tensor_cores=tensor_cores).predict_sum(x, S)
    assert got.shape == (8, 10)
    np.testing.assert_allclose(got.cpu().numpy(), want.cpu().numpy(), rtol=0, atol=2e-6)    # same ints; fp32 sum order only
    one = Int8MCEngine(net, chunk=1, tensor_cores=tensor_cores)
    np.testing.assert_allclose(one.predict_sum(x, S).cpu().numpy(), got.cpu().numpy(), rtol=0, atol=2e-6)
    tail = Int8MCEngine(net, chunk=8, tensor_cores=tensor_cores).predict_sum(x, 2, sample0=4)
    np.testing.assert_allclose(tail.cpu().numpy(), (loop[4] + loop[5]).cpu().numpy(), rtol=0, atol=2e-6)
    p = Int8MCEngine(net, tensor_cores=tensor_cores).predict(x, S)
    np.testing.assert_allclose(p.sum(-1).cpu().numpy(), np.ones(8), atol=1e-5)


def test_int8_engine_full_resnet_tensor_core_and_imad_paths_agree():
    """Full-size int8 ResNet-18 (24/48/96/192 channels, B=64): lifecycle on the device, then the sample-batched engine on the
    tcgen05 kind::i8 kernel and on the CUDA-core integer kernel — integer arithmetic is exact, so the class probabilities
    must be identical, not just close."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "bench_int8.py")
    spec = importlib.util.spec_from_file_location("bench_int8", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    from qbn_b200.mc_int8 import Int8MCEngine
    net, x, _ = mod.build_model(B=64)
    fast = Int8MCEngine(net, chunk=3, tensor_cores=True).predict(x, 3)
    slow = Int8MCEngine(net, chunk=3, tensor_cores=False).predict(x, 3)
    assert fast.shape == (64, 10) and torch.isfinite(fast).all()
    assert torch.equal(fast, slow)
    assert float((fast.sum(-1) - 1).abs().max()) < 1e-5


def test_lenet_int8_network_with_reference_state_end_to_end(golden):
    """The LeNet-shaped converted net (max-pooling between the convolutions, fused Linear+ReLU) loaded with the reference's
    int8 state and replayed noise reproduces the reference's int8 forward end to end; the sample-batched engine equals the
    per-sample loop.  (Same checks as tests/test_int8_host_logic_cpu.py, here with the CUDA kernels instead of the stand-ins.)"""
    import __graft_entry__ as ge
    ge.build()
    from qbn_b200 import noise, zoo
    from qbn_b200.mc_int8 import Int8MCEngine
    from test_int8_host_logic_cpu import _lenet_int8
    g = golden("tiny_int8")
    args = zoo.Args(sigma_prior=0.1, model="conv_lenet_bbb", q=True, at=True, activation_precision=7, weight_precision=8)
    m = _lenet_int8(g, args, device="cuda").cuda()
    q_names = [str(n) for n in g["q_names"]]
    x = torch.as_tensor(g["x"]).cuda()
    with torch.no_grad(), noise.inject([torch.as_tensor(g[n + ".eps"]).cuda() for n in q_names]):
        y = m(x)
    np.testing.assert_allclose(y.cpu().numpy(), g["y"], rtol=1e-5, atol=1e-7)
    noise.manual_seed(4)
    with torch.no_grad():
        loop = []
        for s in range(3):
            with noise.sample_index(s):
                loop.append(m(x))
    got = Int8MCEngine(m, chunk=3).predict_sum(x, 3)
    np.testing.assert_allclose(got.cpu().numpy(), torch.stack(loop).sum(0).cpu().numpy(), rtol=0, atol=2e-6)
