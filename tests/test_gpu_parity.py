"""GPU parity tests (pytest -m gpu): the CUDA path, called through the C ABI (qbn_b200.ops ->
ctypes -> libqbn.so), against the oracle and the reference-generated golden vectors.
fp32 mode: rtol 1e-5 (+ atol 1e-5 * max|ref| for near-zero outputs); integer paths: bit-exact."""
import numpy as np
import pytest
import torch

import oracle.philox as OP
import oracle.qbn_oracle as O

pytestmark = pytest.mark.gpu


def dev(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def close(got, ref, rtol=1e-5, atol_rel=1e-5):
    got = got.detach().cpu().numpy() if torch.is_tensor(got) else np.asarray(got)
    ref = ref.detach().cpu().numpy() if torch.is_tensor(ref) else np.asarray(ref)
    atol = atol_rel * max(1e-30, float(np.abs(ref).max()))
    np.testing.assert_allclose(got, ref, rtol=rtol, atol=atol)


@pytest.fixture(scope="module", autouse=True)
def _lib():
    import __graft_entry__ as g
    g.build()
    from qbn_b200 import config
    config.set_math_mode("fp32")


# ------------------------------------------------------------------------------------------------
def test_device_is_blackwell():
    import ctypes
    from qbn_b200 import _lib as L
    sm, ma, mi = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    L.call("qbn_device_info", ctypes.byref(sm), ctypes.byref(ma), ctypes.byref(mi))
    assert ma.value == 10, "libqbn is built for sm_100a only"


def test_philox_bits_and_moments():
    from qbn_b200 import ops
    n = 100003
    got = ops.philox_u32(n, 0x1234567890ABCDEF, 7, 3).cpu().numpy().view(np.uint32)
    ref = OP.philox_u32(n, 0x1234567890ABCDEF, 7, 3)
    assert np.array_equal(got, ref)
    z = ops.philox_normal(n, 42, 1, 2).cpu().numpy()
    zr = OP.philox_normal(n, 42, 1, 2)
    np.testing.assert_allclose(z, zr, rtol=0, atol=2e-5)
    big = ops.philox_normal(4_000_000, 99, 5, 0).double()
    assert abs(float(big.mean())) < 4 * 1.0 / np.sqrt(4e6)
    assert abs(float(big.var()) - 1.0) < 4 * np.sqrt(2.0 / 4e6)
    assert abs(float((big ** 3).mean())) < 0.01            # skewness
    assert abs(float((big ** 4).mean()) - 3.0) < 0.03      # kurtosis
    # Kolmogorov-Smirnov against N(0,1)
    from scipy import stats
    ks = stats.kstest(big[:200000].cpu().numpy(), "norm")
    assert ks.pvalue > 1e-3
    # independent substreams: different (layer, sample) ids are uncorrelated
    a = ops.philox_normal(1_000_000, 99, 5, 0)
    b = ops.philox_normal(1_000_000, 99, 5, 1)
    assert abs(float((a * b).mean())) < 5e-3
    m = ops.philox_bernoulli(2_000_000, 0.85, 7, 0, 0)
    assert abs(float(m.mean()) - 0.85) < 4 * np.sqrt(0.85 * 0.15 / 2e6)
    assert np.array_equal(ops.philox_bernoulli(1000, 0.8, 7, 1, 2).cpu().numpy(), OP.philox_bernoulli(1000, 0.8, 7, 1, 2))


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_linear_golden(golden, tag):
    from qbn_b200 import noise
    from qbn_b200.stochastic.bbb.linear import Linear
    g = golden("linear_" + tag)
    has_bias = bool(g["has_bias"])
    N, K = g["mu"].shape
    lin = Linear(K, N, has_bias, sigma_prior=float(g["sigma_prior"][0])).cuda()
    with torch.no_grad():
        lin.weight.copy_(dev(g["mu"]))
        lin.std.copy_(dev(g["rho"]))
        if has_bias:
            lin.bias.copy_(dev(g["bias"]))
    x = dev(g["x"]).requires_grad_(True)
    lin.train()
    with noise.inject([dev(g["eps"])]):
        y = lin(x)
    close(y, g["y_train"])
    y.backward(dev(g["gout"]))
    close(x.grad, g["dx"], 1e-4, 1e-5)
    close(lin.weight.grad, g["dmu"], 1e-4, 1e-5)
    close(lin.std.grad, g["drho"], 1e-4, 1e-5)
    if has_bias:
        close(lin.bias.grad, g["dbias"], 1e-4, 1e-5)
    lin.eval()
    with noise.inject([dev(g["eps_w"])]):
        ye = lin(x)
    close(ye, g["y_eval"])
    close(lin.get_kl_divergence(), g["kl"], 1e-5, 1e-6)


@pytest.mark.parametrize("tag", ["a", "b", "c", "d", "e"])
def test_conv_golden(golden, tag):
    from qbn_b200 import noise
    from qbn_b200.stochastic.bbb.conv import Conv2d
    g = golden("conv_" + tag)
    has_bias = bool(g["has_bias"])
    N, C, R, S = g["mu"].shape
    s, p = int(g["stride"]), int(g["pad"])
    conv = Conv2d(C, N, (R, S), stride=s, padding=p, bias=has_bias, sigma_prior=float(g["sigma_prior"][0])).cuda()
    with torch.no_grad():
        conv.weight.copy_(dev(g["mu"]))
        conv.std.copy_(dev(g["rho"]))
        if has_bias:
            conv.bias.copy_(dev(g["bias"]))
    x = dev(g["x"]).requires_grad_(True)
    conv.train()
    saved = conv.bias
    conv.bias = None  # the golden LRT pass was generated without bias (reference quirk, see make_golden.py)
    with noise.inject([dev(g["eps"])]):
        y = conv(x)
    close(y, g["y_train"])
    y.backward(dev(g["gout"]))
    conv.bias = saved
    close(x.grad, g["dx"], 1e-4, 1e-5)
    close(conv.weight.grad, g["dmu"], 1e-4, 1e-5)
    close(conv.std.grad, g["drho"], 1e-4, 1e-5)
    conv.eval()
    with noise.inject([dev(g["eps_w"])]):
        ye = conv(x)
    close(ye, g["y_eval"])
    close(conv.get_kl_divergence(), g["kl"], 1e-5, 1e-6)


def test_lrt_bias_and_philox_backward_consistency():
    """LRT with Philox noise: backward regenerates the same eps as forward (finite-difference free
    check: run forward twice with the same key -> identical; grads match oracle given recovered eps)."""
    from qbn_b200 import ops
    torch.manual_seed(0)
    B, C, H, N = 3, 8, 6, 12
    x = torch.randn(B, C, H, H, device="cuda")
    mu = torch.randn(N, C, 3, 3, device="cuda") * 0.1
    rho = torch.full((N, C, 3, 3), -3.0, device="cuda")
    bias = torch.randn(N, device="cuda")
    key = (1234, 5, 9)
    xr = x.clone().requires_grad_(True)
    mur, rhor, br = mu.clone().requires_grad_(True), rho.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    y = ops.LRTFunction.apply(xr, mur, rhor, br, 1, 1, 1, None, key, 0, False, None)
    y2 = ops.LRTFunction.apply(x, mu, rho, bias, 1, 1, 1, None, key, 0, False, None)
    assert torch.equal(y, y2)
    # recover eps from the forward: eps = (y - mean - bias)/std, compare backward with the oracle's
    mean, std = O.lrt_conv_fwd(x.cpu(), mu.cpu(), rho.cpu(), None, torch.zeros_like(y.cpu()), 1, 1)
    eps = (y.detach().cpu() - mean - bias.cpu().view(1, -1, 1, 1)) / std
    assert abs(float(eps.mean())) < 0.15 and abs(float(eps.std()) - 1) < 0.15
    gout = torch.randn_like(y)
    y.backward(gout)
    dx, dmu, drho, db = O.lrt_conv_bwd(x.cpu(), mu.cpu(), rho.cpu(), eps, std, gout.cpu(), 1, 1)
    close(xr.grad, dx, 1e-3, 1e-4)
    close(mur.grad, dmu, 1e-3, 1e-4)
    close(rhor.grad, drho, 1e-3, 1e-4)
    close(br.grad, db, 1e-4, 1e-5)


def test_kl_grad(golden):
    from qbn_b200 import ops
    g = golden("linear_b")
    mu = dev(g["mu"]).requires_grad_(True)
    rho = dev(g["rho"]).requires_grad_(True)
    kl = ops.kl_divergence(mu, rho, 0.7)
    (kl * 0.25).backward()
    dmu, drho = O.kl_grads(g["mu"], g["rho"], 0.7)
    close(mu.grad, 0.25 * dmu, 1e-5, 1e-6)
    close(rho.grad, 0.25 * drho, 1e-4, 1e-5)
    close(kl, O.kl_divergence(g["mu"], g["rho"], 0.7), 1e-5, 1e-6)


def test_dropout(golden):
    from qbn_b200 import noise
    from qbn_b200.stochastic.mcdropout.dropout import BernoulliDropout
    g = golden("dropout")
    d = BernoulliDropout(float(g["p"])).cuda()
    with noise.inject([dev(g["m4"])]):
        close(d(dev(g["x4"])), g["y4"], 1e-6, 0)
    with noise.inject([dev(g["m2"])]):
        close(d(dev(g["x2"])), g["y2"], 1e-6, 0)
    # Philox mask: keep-rate and scaling
    x = torch.ones(512, 64, 4, 4, device="cuda")
    y = d(x)
    kept = (y[:, :, 0, 0] != 0).float().mean().item()
    assert abs(kept - 0.8) < 0.01
    assert torch.all((y == 0) | (torch.abs(y - 1.25) < 1e-6))
    assert torch.equal(y[:, :, 0, 0], y[:, :, 3, 2])  # one draw per (n, c), broadcast over H, W


def test_metrics_and_mc(golden):
    from qbn_b200 import ops
    g = golden("metrics")
    out = torch.zeros(4 + 30, device="cuda")
    ops.cls_metrics_accumulate(dev(g["probs"]), dev(g["target"]), out)
    o = out.cpu().numpy().astype(np.float64)
    B = g["probs"].shape[0]
    close(o[0] / B, g["error"])
    close(o[1] / B, g["nll"], 1e-5)
    close(o[2] / B, g["brier"], 1e-5)
    close(o[3] / B, g["entropy"], 1e-5)
    bins = o[4:].reshape(10, 3)
    close(O.ece_from_bins(bins), g["ece"], 1e-4, 1e-5)
    ref = O.cls_metric_sums(g["probs"], g["target"])
    assert np.array_equal(bins[:, 2], ref["bins"][:, 2]) and np.array_equal(bins[:, 1], ref["bins"][:, 1])
    close(ops.mc_mean(dev(g["plist"])), g["pmean"], 1e-6, 1e-7)
    mean, var = ops.reg_mc_reduce(dev(g["mus"]), dev(g["vars"]))
    close(mean, g["reg_mean"], 1e-5, 1e-6)
    close(var, g["reg_var"], 1e-5, 1e-6)
    ro = torch.zeros(3, device="cuda")
    ops.reg_metrics_accumulate(mean, var, dev(g["reg_target"]), ro)
    R = g["reg_target"].shape[0]
    r = ro.cpu().numpy().astype(np.float64)
    close(r[0] / R, g["reg_nll"], 1e-5)
    close(r[1] / R, g["reg_mse"], 1e-5)
    close(np.sqrt(r[1] / R), g["reg_rmse"], 1e-5)
    close(r[2] / R, g["reg_mae"], 1e-5)
    # softmax accumulate == sum of softmaxes
    logits = torch.randn(7, 33, 10, device="cuda") * 3
    ps = ops.softmax_accumulate(logits)
    close(ps, torch.softmax(logits, -1).sum(0), 1e-5, 1e-6)
    ps2 = ops.softmax_accumulate(logits, ps.clone())
    close(ps2, 2 * torch.softmax(logits, -1).sum(0), 1e-5, 1e-6)


# ------------------------------------------------------------------------------------------------
# integer path: bit exact
# ------------------------------------------------------------------------------------------------
def _i8_params(s_mu, z_mu, s_sg, z_sg, s_mul, z_mul, s_add, z_add, w_bits=8, n_vec=-1):
    from qbn_b200 import ops
    p = ops.I8SampleParams()
    p.s_mu, p.z_mu, p.s_sigma, p.z_sigma = float(s_mu), int(z_mu), float(s_sg), int(z_sg)
    p.s_eps, p.z_eps = float(O.NOISE_SCALE), 0
    p.s_mul, p.z_mul, p.s_add, p.z_add = float(s_mul), int(z_mul), float(s_add), int(z_add)
    p.w_min, p.w_max = O.INT_BOUNDS[w_bits]
    p.n_vec = n_vec
    return p


def test_quant_ops_bit_exact(golden):
    from qbn_b200 import ops
    g = golden("quant_ops")
    s_mu, z_mu, s_sg, z_sg, s_mul, z_mul, s_add, z_add = g["qp"]
    eq = ops.quantize_s8(dev(g["eps"]), O.NOISE_SCALE, 0)
    assert np.array_equal(eq.cpu().numpy(), g["eps_q"])
    w = ops.i8_sample_weights(dev(g["mu_i"]), dev(g["sg_i"]), _i8_params(s_mu, z_mu, s_sg, z_sg, s_mul, z_mul, s_add, z_add),
                              1, dev(g["eps"]).reshape(1, -1))
    assert np.array_equal(w[0].cpu().numpy(), g["w"])
    # quantized linear (IMAD path: K=70 is ragged)
    s_x, z_x, s_w, z_w, s_o, z_o = g["lin_qp"]
    B, K = g["lin_x"].shape
    N = g["lin_w"].shape[0]
    d = ops.make_desc(B, 1, 1, K, N, 1, 1)
    for relu in (0, 1):
        for hb in (0, 1):
            y, acc = ops.i8_conv_forward(dev(g["lin_x"]), s_x, z_x, dev(g["lin_w"]).reshape(1, -1), s_w, z_w, d,
                                         dev(g["lin_bias"]) if hb else None, s_o, z_o, relu, act_bits=8, want_acc=True, path=1, linear=True)
            assert np.array_equal(y.cpu().numpy(), g["lin_y_relu%d_bias%d" % (relu, hb)]), (relu, hb)
            _, acc_ref = O.i8_linear(g["lin_x"], s_x, int(z_x), g["lin_w"], s_w, int(z_w), None, s_o, int(z_o))
            assert np.array_equal(acc.cpu().numpy(), acc_ref)
    # quantized conv
    Bc, C, H, _ = g["conv_x"].shape
    Nc = g["conv_w"].shape[0]
    xq = dev(g["conv_x"]).contiguous(memory_format=torch.channels_last)
    wq = dev(g["conv_w"]).permute(0, 2, 3, 1).contiguous().reshape(1, -1)
    for stride in (1, 2):
        d = ops.make_desc(Bc, H, H, C, Nc, 3, 3, stride, 1, 1)
        for relu in (0, 1):
            y = ops.i8_conv_forward(xq, s_x, z_x, wq, s_w, z_w, d, dev(g["conv_bias"]), s_o, z_o, relu, act_bits=8, path=1)
            assert np.array_equal(y.cpu().numpy(), g["conv_y_s%d_relu%d" % (stride, relu)]), (stride, relu)
    sa, za, sb, zb, so, zo = g["add_qp"]
    y = ops.i8_add(dev(g["add_a"]), sa, za, dev(g["add_b"]), sb, zb, so, zo, act_bits=8)
    assert np.array_equal(y.cpu().numpy(), g["add_y"])
    # fake quantise (given qparams: observe=False)
    s, z, qmin, qmax = g["fq_qp"]
    fq = ops.FakeQuantState(int(qmin), int(qmax))
    fq.scale.fill_(float(s))
    fq.zero_point.fill_(int(z))
    xf = dev(g["fq_x"]).requires_grad_(True)
    yf = ops.fake_quantize(xf, fq, observe=False)
    assert np.array_equal(yf.detach().cpu().numpy(), g["fq_y"])
    yf.backward(torch.ones_like(yf))
    assert np.array_equal(xf.grad.cpu().numpy(), g["fq_mask"])
    # int8 dropout
    s_x, z_x, s_m, z_m, mult = g["do_qp"]
    xq = dev(g["do_x"]).contiguous(memory_format=torch.channels_last)
    y = ops.i8_dropout(xq, s_x, z_x, 0.15, s_m, z_m, dev(g["do_mask"]), act_bits=8)
    assert np.array_equal(y.cpu().numpy(), g["do_y"])


def test_observer_matches_torch():
    """MovingAverageMinMaxObserver + qparams: first call initialises, later calls EMA (c=0.01)."""
    from torch.ao.quantization import FakeQuantize, MovingAverageMinMaxObserver
    from qbn_b200 import ops
    ref = FakeQuantize(observer=MovingAverageMinMaxObserver, quant_min=-128, quant_max=127, dtype=torch.qint8,
                       qscheme=torch.per_tensor_affine)
    fq = ops.FakeQuantState(-128, 127)
    g = torch.Generator().manual_seed(5)
    for it in range(4):
        x = torch.randn(5000, generator=g) * (0.5 + it) + 0.1 * it
        yr = ref(x)
        y = ops.fake_quantize(x.cuda(), fq, observe=True)
        close(fq.scale, ref.scale, 1e-6, 0)
        assert int(fq.zero_point.item()) == int(ref.zero_point.item())
        close(fq.state[:2], torch.stack([ref.activation_post_process.min_val, ref.activation_post_process.max_val]), 1e-6, 0)
        close(y, yr, 1e-6, 1e-7)


def test_tiny_int8_layers_bit_exact(golden):
    """Per-layer int8 forward of the converted LeNet-shaped net: sampled weights and outputs equal the
    reference's FBGEMM path bit for bit (both the IMAD kernel and, where aligned, tcgen05 kind::i8)."""
    from qbn_b200 import ops
    g = golden("tiny_int8")
    for n in g["q_names"]:
        s_mu, z_mu = g[n + ".mu_qp"]
        s_sg, z_sg = g[n + ".sigma_qp"]
        s_mul, z_mul = g[n + ".mul_qp"]
        s_add, z_add = g[n + ".add_qp"]
        s_x, z_x = g[n + ".x_qp"]
        s_o, z_o = g[n + ".out_qp"]
        relu = bool(g[n + ".relu"])
        mu_q, sg_q, eps, w_ref = g[n + ".mu_q"], g[n + ".sigma_q"], g[n + ".eps"], g[n + ".w_q"]
        is_conv = mu_q.ndim == 4
        # the reference samples in OIHW order: vector/tail split of ATen's qadd refers to that order,
        # so sample in OIHW and pack afterwards
        params = _i8_params(s_mu, z_mu, s_sg, z_sg, s_mul, z_mul, s_add, z_add, 8)
        w = ops.i8_sample_weights(dev(mu_q).reshape(-1), dev(sg_q).reshape(-1), params, 1, dev(eps).reshape(1, -1))
        assert np.array_equal(w[0].cpu().numpy().reshape(w_ref.shape), w_ref), n
        x_q = g[n + ".x_q"]
        if is_conv:
            stride, pad = [int(v) for v in g[n + ".conv"]]
            N, C, R, S = mu_q.shape
            B, _, H, W = x_q.shape
            d = ops.make_desc(B, H, W, C, N, R, S, stride, pad, 1)
            wp = w[0].reshape(N, C, R, S).permute(0, 2, 3, 1).contiguous().reshape(1, -1)
            xq = dev(x_q).contiguous(memory_format=torch.channels_last)
            y = ops.i8_conv_forward(xq, s_x, z_x, wp, s_add, z_add, d, None, s_o, z_o, relu, act_bits=8, path=1)  # module output, before clamp_activation
        else:
            N, K = mu_q.shape
            d = ops.make_desc(x_q.shape[0], 1, 1, K, N, 1, 1)
            y = ops.i8_conv_forward(dev(x_q), s_x, z_x, w, s_add, z_add, d, None, s_o, z_o, relu, act_bits=8, path=1, linear=True)
        assert np.array_equal(y.cpu().numpy(), g[n + ".y_q"]), n


def test_sample_weights_philox_matches_oracle_stream():
    """Philox eval sampling: W = mu + sigma*eps with eps = oracle Philox stream (layer, sample)."""
    from qbn_b200 import ops
    n, S = 1000, 3
    mu = torch.randn(n, device="cuda")
    sg = torch.rand(n, device="cuda") * 0.1
    w = ops.sample_weights(mu, sg, S, None, seed=77, layer_id=4, sample0=10)
    for s in range(S):
        eps = torch.as_tensor(OP.philox_normal(n, 77, 4, 10 + s)).cuda()
        close(w[s], mu + sg * eps, 1e-5, 1e-6)
    # sharding independence: samples 11..12 drawn alone equal rows 1..2 of the full draw
    w2 = ops.sample_weights(mu, sg, 2, None, seed=77, layer_id=4, sample0=11)
    assert torch.equal(w2, w[1:])


def test_kl_multi_matches_per_layer_sum():
    """qbn_kl_multi (all layers, one launch) == the sum of the per-layer KL terms, values and gradients."""
    from qbn_b200 import ops
    g = torch.Generator().manual_seed(17)
    shapes = [(24, 3, 3, 3), (48, 24, 3, 3), (10, 192), (7,)]
    priors = [0.05, 0.05, 0.1, 1.0]
    ps = [(torch.randn(s, generator=g).cuda().requires_grad_(True), (torch.randn(s, generator=g) - 3).cuda().requires_grad_(True)) for s in shapes]
    ref = sum(ops.kl_divergence(mu, rho, sp) for (mu, rho), sp in zip(ps, priors))
    (ref * 0.37).backward()
    ref_g = [(mu.grad.clone(), rho.grad.clone()) for mu, rho in ps]
    for mu, rho in ps:
        mu.grad = None
        rho.grad = None
    for _ in range(2):                                   # second call reuses the cached job table
        got = ops.kl_divergence_multi([(mu, rho, sp) for (mu, rho), sp in zip(ps, priors)])
        (got * 0.37).backward()
        np.testing.assert_allclose(float(got.detach()), float(ref.detach()), rtol=1e-6)
        for (mu, rho), (gm, gr) in zip(ps, ref_g):
            np.testing.assert_allclose(mu.grad.cpu().numpy(), gm.cpu().numpy(), rtol=1e-5, atol=1e-8)
            np.testing.assert_allclose(rho.grad.cpu().numpy(), gr.cpu().numpy(), rtol=1e-5, atol=1e-8)
            mu.grad = None
            rho.grad = None
