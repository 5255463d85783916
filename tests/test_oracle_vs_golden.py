"""CPU: pins oracle/qbn_oracle.py (the restatement) against golden vectors produced by the
UNMODIFIED reference (oracle/make_golden.py).  Float functions: rtol 1e-5 (fp32 re-association
only); integer functions: bit-exact."""
import numpy as np
import pytest
import torch

import oracle.qbn_oracle as O

RT, AT = 1e-5, 1e-6


def close(a, b, rtol=RT, atol=AT):
    np.testing.assert_allclose(np.asarray(a), np.asarray(b), rtol=rtol, atol=atol)


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_linear(golden, tag):
    g = golden("linear_" + tag)
    bias = g["bias"] if bool(g["has_bias"]) else None
    y, std = O.lrt_linear_fwd(g["x"], g["mu"], g["rho"], bias, g["eps"])
    close(y, g["y_train"])
    dx, dmu, drho, db = O.lrt_linear_bwd(g["x"], g["mu"], g["rho"], g["eps"], std, g["gout"])
    close(dx, g["dx"], 1e-4, 1e-5)
    close(dmu, g["dmu"], 1e-4, 1e-5)
    close(drho, g["drho"], 1e-4, 1e-6)
    if bias is not None:
        close(db, g["dbias"], 1e-4, 1e-5)
    close(O.eval_linear_fwd(g["x"], g["mu"], g["rho"], bias, g["eps_w"]), g["y_eval"])
    close(O.kl_divergence(g["mu"], g["rho"], float(g["sigma_prior"][0])), g["kl"], 1e-5, 1e-3)


@pytest.mark.parametrize("tag", ["a", "b", "c", "d", "e"])
def test_conv(golden, tag):
    g = golden("conv_" + tag)
    s, p = int(g["stride"]), int(g["pad"])
    y, std = O.lrt_conv_fwd(g["x"], g["mu"], g["rho"], None, g["eps"], s, p)
    close(y, g["y_train"])
    dx, dmu, drho, _ = O.lrt_conv_bwd(g["x"], g["mu"], g["rho"], g["eps"], std, g["gout"], s, p)
    close(dx, g["dx"], 1e-4, 1e-5)
    close(dmu, g["dmu"], 1e-4, 1e-4)
    close(drho, g["drho"], 1e-4, 1e-5)
    bias = g["bias"] if bool(g["has_bias"]) else None
    close(O.eval_conv_fwd(g["x"], g["mu"], g["rho"], bias, g["eps_w"], s, p), g["y_eval"])
    close(O.kl_divergence(g["mu"], g["rho"], float(g["sigma_prior"][0])), g["kl"], 1e-5, 1e-3)


def test_kl_grads_match_autograd(golden):
    g = golden("linear_a")
    mu = torch.tensor(g["mu"], requires_grad=True)
    rho = torch.tensor(g["rho"], requires_grad=True)
    O.kl_divergence(mu, rho, 0.7).backward()
    dmu, drho = O.kl_grads(g["mu"], g["rho"], 0.7)
    close(dmu, mu.grad, 1e-5, 1e-6)
    close(drho, rho.grad, 1e-4, 1e-5)


def test_dropout(golden):
    g = golden("dropout")
    close(O.dropout_fwd(g["x4"], g["m4"], float(g["p"])), g["y4"])
    close(O.dropout_fwd(g["x2"], g["m2"], float(g["p"])), g["y2"])


def test_metrics(golden):
    g = golden("metrics")
    m = O.cls_metric_sums(g["probs"], g["target"])
    B = g["probs"].shape[0]
    close(m["error"] / B, g["error"])
    close(m["nll"] / B, g["nll"])
    close(m["brier"] / B, g["brier"])
    close(m["entropy"] / B, g["entropy"])
    close(O.ece_from_bins(m["bins"]), g["ece"], 1e-5, 1e-6)
    mean, var = O.reg_mc_reduce(list(g["mus"]), list(g["vars"]))
    close(mean, g["reg_mean"])
    close(var, g["reg_var"])
    r = O.reg_metric_sums(mean, var, g["reg_target"])
    R = g["reg_target"].shape[0]
    close(r["nll"] / R, g["reg_nll"])
    close(r["se"] / R, g["reg_mse"])
    close(np.sqrt(r["se"] / R), g["reg_rmse"])
    close(r["ae"] / R, g["reg_mae"])
    close(O.mc_mean_probs(list(g["plist"])), g["pmean"])


def test_quant_ops_bit_exact(golden):
    g = golden("quant_ops")
    s_mu, z_mu, s_sig, z_sig, s_mul, z_mul, s_add, z_add = g["qp"]
    eq = O.quantize(g["eps"], O.NOISE_SCALE, 0, -128, 127)
    assert np.array_equal(eq, g["eps_q"])
    r = O.qmul(g["sg_i"], s_sig, int(z_sig), eq, O.NOISE_SCALE, 0, s_mul, int(z_mul))
    assert np.array_equal(r, g["r"])
    w = O.qadd(g["mu_i"], s_mu, int(z_mu), r, s_mul, int(z_mul), s_add, int(z_add))
    assert np.array_equal(w, g["w"])
    w2 = O.i8_sample_weight(g["mu_i"], s_mu, int(z_mu), g["sg_i"], s_sig, int(z_sig), g["eps"], s_mul, int(z_mul), s_add, int(z_add))
    assert np.array_equal(w2, g["w"])
    s_x, z_x, s_w, z_w, s_o, z_o = g["lin_qp"]
    for relu in (0, 1):
        for hb in (0, 1):
            y, _ = O.i8_linear(g["lin_x"], s_x, int(z_x), g["lin_w"], s_w, int(z_w), g["lin_bias"] if hb else None, s_o, int(z_o), bool(relu))
            assert np.array_equal(y, g["lin_y_relu%d_bias%d" % (relu, hb)]), (relu, hb)
    for stride in (1, 2):
        for relu in (0, 1):
            y, _ = O.i8_conv(g["conv_x"], s_x, int(z_x), g["conv_w"], s_w, int(z_w), g["conv_bias"], s_o, int(z_o), stride, 1, 1, bool(relu))
            assert np.array_equal(y, g["conv_y_s%d_relu%d" % (stride, relu)]), (stride, relu)
    sa, za, sb, zb, so, zo = g["add_qp"]
    assert np.array_equal(O.qadd(g["add_a"], sa, int(za), g["add_b"], sb, int(zb), so, int(zo), 0, 255), g["add_y"])
    s, z, qmin, qmax = g["fq_qp"]
    y, mask = O.fake_quant(g["fq_x"], s, int(z), int(qmin), int(qmax))
    assert np.array_equal(y, g["fq_y"])
    assert np.array_equal(mask.astype(np.float32), g["fq_mask"])
    s_x, z_x, s_m, z_m, mult = g["do_qp"]
    q, s_new, z_new = O.i8_dropout(g["do_x"], s_x, int(z_x), g["do_mask"], s_m, int(z_m), mult)
    assert np.array_equal(q, g["do_y"])
    close(s_new, g["do_y_qp"][0], 1e-6, 0)
    assert z_new == int(g["do_y_qp"][1])


def test_tiny_int8_layers_bit_exact(golden):
    g = golden("tiny_int8")
    for n in g["q_names"]:
        s_mu, z_mu = g[n + ".mu_qp"]
        s_sg, z_sg = g[n + ".sigma_qp"]
        s_mul, z_mul = g[n + ".mul_qp"]
        s_add, z_add = g[n + ".add_qp"]
        s_x, z_x = g[n + ".x_qp"]
        s_o, z_o = g[n + ".out_qp"]
        w = O.i8_sample_weight(g[n + ".mu_q"], s_mu, int(z_mu), g[n + ".sigma_q"], s_sg, int(z_sg), g[n + ".eps"],
                               s_mul, int(z_mul), s_add, int(z_add), w_bits=8)
        assert np.array_equal(w, g[n + ".w_q"]), n
        relu = bool(g[n + ".relu"])
        if w.ndim == 4:
            stride, pad = g[n + ".conv"]
            y, _ = O.i8_conv(g[n + ".x_q"], s_x, int(z_x), w, s_add, int(z_add), None, s_o, int(z_o), int(stride), int(pad), 1, relu, act_bits=8)
        else:
            y, _ = O.i8_linear(g[n + ".x_q"], s_x, int(z_x), w, s_add, int(z_add), None, s_o, int(z_o), relu, act_bits=8)
        assert np.array_equal(y, g[n + ".y_q"]), n


def _i8_layer(g, n, x_q=None, x_qp=None, act_bits=8):
    """One converted BBB layer of an int8 fixture through the oracle: (sampled int8 weight, integer output)."""
    qp = lambda k: (float(g[n + k][0]), int(g[n + k][1]))
    (s_mu, z_mu), (s_sg, z_sg), (s_mul, z_mul), (s_add, z_add), (s_o, z_o) = qp(".mu_qp"), qp(".sigma_qp"), qp(".mul_qp"), qp(".add_qp"), qp(".out_qp")
    s_x, z_x = qp(".x_qp") if x_qp is None else x_qp
    x_q = g[n + ".x_q"] if x_q is None else x_q
    w = O.i8_sample_weight(g[n + ".mu_q"], s_mu, z_mu, g[n + ".sigma_q"], s_sg, z_sg, g[n + ".eps"], s_mul, z_mul, s_add, z_add, w_bits=8)
    relu = bool(g[n + ".relu"])
    bias = g[n + ".bias"] if (n + ".bias") in g.files and g[n + ".bias"].size else None
    if w.ndim == 4:
        stride, pad = g[n + ".conv"]
        y, _ = O.i8_conv(x_q, s_x, z_x, w, s_add, z_add, bias, s_o, z_o, int(stride), int(pad), 1, relu, act_bits=act_bits)
    else:
        y, _ = O.i8_linear(x_q, s_x, z_x, w, s_add, z_add, bias, s_o, z_o, relu, act_bits=act_bits)
    return w, y, (s_o, z_o)


def test_tiny_resnet_int8_layers_bit_exact(golden):
    """Every captured node of the converted ResNet-shaped net, one at a time from the reference's own integer inputs."""
    g = golden("tiny_resnet_int8")
    for n in [str(v) for v in g["order"]]:
        if n + ".mu_q" in g.files:
            w, y, _ = _i8_layer(g, n)
            assert np.array_equal(w, g[n + ".w_q"]), n
            assert np.array_equal(y, g[n + ".y_q"]), n
        elif n.endswith(".add"):
            (sa, za), (sb, zb), (so, zo) = g[n + ".a_qp"], g[n + ".b_qp"], g[n + ".y_qp"]
            # the reference's operands are channels_last: ATen walks them in NHWC memory order
            a, b = (np.ascontiguousarray(g[n + k].transpose(0, 2, 3, 1)) for k in (".a_q", ".b_q"))
            y = O.i8_add(a, sa, int(za), b, sb, int(zb), so, int(zo), act_bits=8).transpose(0, 3, 1, 2)
            assert np.array_equal(y, g[n + ".y_q"]), n
        elif n + ".add.y_q" in g.files:                       # BasicBlock: checked as a whole by the chained test
            continue
        else:
            assert np.array_equal(O.i8_avgpool(g[n + ".x_q"], int(g[n + ".y_qp"][1]), 4, act_bits=7), g[n + ".y_q"]), n


def test_tiny_resnet_int8_chained_end_to_end(golden):
    """The whole int8 network chained through the oracle from the float input: quantise, stem, identity block,
    stride-2 block with its 1x1 shortcut, quantised ReLU / add / average pool, linear, dequantise, softmax."""
    g = golden("tiny_resnet_int8")
    A = 7                                                         # activation_precision of the fixture
    s, z = float(g["quant_qp"][0]), int(g["quant_qp"][1])
    h = np.clip(O.quantize(g["x"], s, z, 0, 255), *O.UINT_BOUNDS[A])
    _, h, (s, z) = _i8_layer(g, "layers.0", h, (s, z), act_bits=A)
    assert np.array_equal(h, g["layers.0.y_q"])
    for blk in ("layers.3.0", "layers.3.1"):
        assert np.array_equal(h, g[blk + ".x_q"]), blk
        _, t, qp1 = _i8_layer(g, blk + ".stem.0", h, (s, z), act_bits=A)
        _, t, qp2 = _i8_layer(g, blk + ".stem.3", t, qp1, act_bits=A)
        if blk + ".shortcut.0.mu_q" in g.files:
            _, sc, qps = _i8_layer(g, blk + ".shortcut.0", h, (s, z), act_bits=A)
        else:
            sc, qps = h, (s, z)
        so, zo = float(g[blk + ".add.y_qp"][0]), int(g[blk + ".add.y_qp"][1])
        nhwc = lambda v: np.ascontiguousarray(v.transpose(0, 2, 3, 1))
        t = O.i8_add(nhwc(t), qp2[0], qp2[1], nhwc(sc), qps[0], qps[1], so, zo, act_bits=A).transpose(0, 3, 1, 2)
        assert np.array_equal(t, g[blk + ".add.y_q"]), blk
        h, (s, z) = O.i8_relu(t, zo, act_bits=A), (so, zo)
        assert np.array_equal(h, g[blk + ".y_q"]), blk
    h = O.i8_avgpool(h, z, 4, act_bits=A)
    assert np.array_equal(h, g["layers.4.y_q"])
    _, h, (s, z) = _i8_layer(g, "layers.6", h.reshape(h.shape[0], -1), (s, z), act_bits=A)
    assert np.array_equal(h, g["layers.6.y_q"])
    logits = torch.as_tensor(O.dequantize(h, s, z))
    close(torch.softmax(logits, dim=-1), g["y"], 1e-6, 1e-7)


def _eps_fn_from(noise_by_name):
    return lambda name, shape: noise_by_name[name]


def test_resnet_eval(golden):
    g = golden("resnet")
    P = O.ResNetBBBParams(seed=21)
    x = torch.randn(4, 3, 32, 32, generator=torch.Generator().manual_seed(22))
    plan = O.resnet_noise_plan(P)
    for s in range(2):
        noise = dict(zip([p[0] for p in plan], O.replay_noise(700 + s, [p[1] for p in plan])))
        y = O.resnet_bbb_eval_forward(P, x, _eps_fn_from(noise))
        close(y, g["y_eval%d" % s], 1e-4, 1e-6)


def test_lenet_and_mlp_eval(golden):
    g = golden("lenet")
    P = O.LeNetBBBParams(seed=31)
    x = torch.rand(4, 1, 28, 28, generator=torch.Generator().manual_seed(32))
    plan = P.noise_plan()
    noise = dict(zip([p[0] for p in plan], O.replay_noise(800, [p[1] for p in plan])))
    close(O.lenet_bbb_eval_forward(P, x, _eps_fn_from(noise)), g["y_eval0"], 1e-4, 1e-6)
    g = golden("mlp")
    P = O.MLPBBBParams(seed=41)
    x = torch.randn(16, 1, generator=torch.Generator().manual_seed(42))
    plan = P.noise_plan()
    noise = dict(zip([p[0] for p in plan], O.replay_noise(900, [p[1] for p in plan])))
    mu, var = O.mlp_bbb_eval_forward(P, x, _eps_fn_from(noise))
    close(mu, g["y_mu"], 1e-4, 1e-6)
    close(var, g["y_var"], 1e-4, 1e-6)


def test_mc_dropout_networks(golden):
    """models_mc.py ResNet / LeNet restatements + the mask replay order against the reference's seeded forwards."""
    g = golden("resnet_mc")
    P = O.ResNetBBBParams(seed=51)
    x = torch.randn(4, 3, 32, 32, generator=torch.Generator().manual_seed(52))
    plan = O.resnet_mc_mask_plan(P, 4)
    for s in range(2):
        masks = dict(zip([n for n, _ in plan], O.replay_masks(1100 + s, [sh for _, sh in plan], 0.15)))
        y = O.resnet_mc_forward(P, x, lambda name, shape: masks[name], 0.15)
        np.testing.assert_allclose(y.numpy(), g["y%d" % s], rtol=1e-5, atol=1e-7)
    g = golden("lenet_mc")
    P = O.LeNetBBBParams(seed=61)
    x = torch.rand(4, 1, 28, 28, generator=torch.Generator().manual_seed(62))
    shapes = [("layers.1", (4, 20)), ("layers.4", (4, 50)), ("layers.9", (4, 500))]
    for s in range(2):
        masks = dict(zip([n for n, _ in shapes], O.replay_masks(1200 + s, [sh for _, sh in shapes], 0.2)))
        y = O.lenet_mc_forward(P, x, lambda name, shape: masks[name], 0.2)
        np.testing.assert_allclose(y.numpy(), g["y%d" % s], rtol=1e-5, atol=1e-7)


def test_full_resnet_int8_fixture_noise_stream_is_reproducible(golden, golden_dir):
    """tests/golden/resnet_int8_full.npz does not store its 2 x 6.3 MB of noise: the GPU tests redraw it from torch's CPU
    generator with the recorded seeds (oracle/make_golden.py:gen_full_resnet_int8).  Pin that the stream is still the one the
    fixture was made with, and that the checkpoint holds the 21 int8 layers of the full-size network."""
    import hashlib
    import torch
    g = golden("resnet_int8_full")
    sd = torch.load(golden_dir / "resnet_int8_full_weights.pt", map_location="cpu")
    q_names = [str(n) for n in g["q_names"]]
    assert len(q_names) == 21 and sum(int(sd[n + ".weight"].numel()) for n in q_names) == 1571592      # stochastic weights
    for fi in (0, 1):
        torch.manual_seed(int(g["seeds"][fi]))
        h = hashlib.sha1()
        for n in q_names:
            h.update(np.ascontiguousarray(torch.empty(tuple(sd[n + ".std"].shape)).normal_().numpy()).tobytes())
        assert h.hexdigest() == str(g["f%d.eps_sha1" % fi])
