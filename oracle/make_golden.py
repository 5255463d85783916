"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by executing the UNMODIFIED reference
modules from /root/reference (CPU, this container only) on seeded inputs.

Noise replay: every stochastic draw of the reference goes through torch's global
CPU generator in forward order as a contiguous tensor, so

    torch.manual_seed(s); y = layer(x)            # reference draws eps internally
    torch.manual_seed(s); eps = torch.empty(shape).normal_()   # identical eps, replayed

The fixtures store x, parameters, the replayed noise and the reference outputs; the parity tests
hand the same noise to the CUDA kernels' "injected noise" arguments.

Run:  python oracle/make_golden.py        (re-creates every fixture deterministically)
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.ref_harness import Args, import_reference  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def npy(t):
    return t.detach().cpu().numpy().copy()


def save(name, **arrays):
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote %-28s %7.1f KB" % (name + ".npz", os.path.getsize(path) / 1024))


def trained_like_(mod, g):
    """Move a freshly constructed BBB layer to trained-like ranges so sigma matters."""
    with torch.no_grad():
        fan_in = mod.weight[0].numel()
        mod.weight.copy_(torch.randn(mod.weight.shape, generator=g) / fan_in ** 0.5)
        mod.std.copy_(torch.empty(mod.std.shape).uniform_(-5.0, -1.0, generator=g))
        if mod.bias is not None:
            mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.1)


def gen_linear(src):
    from src.models.stochastic.bbb.linear import Linear
    g = torch.Generator().manual_seed(11)
    for tag, (B, K, N, bias) in {"a": (13, 37, 19, True), "b": (64, 100, 100, True), "c": (5, 1, 100, True), "d": (9, 100, 1, False)}.items():
        lin = Linear(K, N, bias, sigma_prior=0.7)
        trained_like_(lin, g)
        x = torch.randn(B, K, generator=g).requires_grad_(True)
        gout = torch.randn(B, N, generator=g)
        # train (LRT) forward/backward
        lin.train()
        torch.manual_seed(100)
        y = lin(x)
        y.backward(gout)
        torch.manual_seed(100)
        eps = torch.empty(B, N).normal_()
        # eval (weight sampling) forward
        lin.eval()
        torch.manual_seed(200)
        with torch.no_grad():
            ye = lin(x)
        torch.manual_seed(200)
        eps_w = torch.empty(N, K).normal_()
        kl = lin.get_kl_divergence()
        save("linear_" + tag, x=npy(x), mu=npy(lin.weight), rho=npy(lin.std),
             bias=npy(lin.bias) if bias else np.zeros(0, np.float32), has_bias=np.array(bias),
             sigma_prior=npy(lin.std_prior), eps=npy(eps), gout=npy(gout), y_train=npy(y),
             dx=npy(x.grad), dmu=npy(lin.weight.grad), drho=npy(lin.std.grad),
             dbias=npy(lin.bias.grad) if bias else np.zeros(0, np.float32),
             eps_w=npy(eps_w), y_eval=npy(ye), kl=npy(kl))


def gen_conv(src):
    from src.models.stochastic.bbb.conv import Conv2d
    g = torch.Generator().manual_seed(12)
    cases = {
        "a": dict(B=3, C=5, N=7, H=9, W=9, k=3, stride=1, pad=1, bias=False),
        "b": dict(B=2, C=4, N=6, H=10, W=8, k=3, stride=2, pad=1, bias=True),
        "c": dict(B=2, C=1, N=3, H=12, W=12, k=5, stride=1, pad=2, bias=False),
        "d": dict(B=2, C=8, N=16, H=8, W=8, k=1, stride=2, pad=0, bias=False),
        "e": dict(B=2, C=24, N=24, H=8, W=8, k=3, stride=1, pad=1, bias=False),
    }
    for tag, c in cases.items():
        conv = Conv2d(c["C"], c["N"], (c["k"], c["k"]), stride=c["stride"], padding=c["pad"], bias=c["bias"], sigma_prior=0.05)
        trained_like_(conv, g)
        x = torch.randn(c["B"], c["C"], c["H"], c["W"], generator=g).requires_grad_(True)
        # NOTE reference quirk: conv.py:32 adds a [N] bias to an NCHW tensor without reshaping, so the
        # LRT (train) branch only broadcasts when Wo == N; every reference model uses bias=False.
        # The train-mode fixture is therefore generated with the bias detached (bias case: eval only).
        conv.train()
        saved_bias = conv.bias
        conv.bias = None
        torch.manual_seed(300)
        y = conv(x)
        gout = torch.randn(y.shape, generator=g)
        y.backward(gout)
        torch.manual_seed(300)
        eps = torch.empty(y.shape).normal_()
        conv.bias = saved_bias
        conv.eval()
        torch.manual_seed(400)
        with torch.no_grad():
            ye = conv(x)
        torch.manual_seed(400)
        eps_w = torch.empty(conv.weight.shape).normal_()
        kl = conv.get_kl_divergence()
        save("conv_" + tag, x=npy(x), mu=npy(conv.weight), rho=npy(conv.std),
             bias=npy(conv.bias) if c["bias"] else np.zeros(0, np.float32), has_bias=np.array(c["bias"]),
             stride=np.array(c["stride"]), pad=np.array(c["pad"]), sigma_prior=npy(conv.std_prior.float()),
             eps=npy(eps), gout=npy(gout), y_train=npy(y), dx=npy(x.grad), dmu=npy(conv.weight.grad),
             drho=npy(conv.std.grad),
             eps_w=npy(eps_w), y_eval=npy(ye), kl=npy(kl))


def gen_dropout(src):
    from src.models.stochastic.mcdropout.dropout import BernoulliDropout
    g = torch.Generator().manual_seed(13)
    d = BernoulliDropout(0.2)
    x4 = torch.randn(6, 10, 5, 5, generator=g)
    x2 = torch.randn(16, 50, generator=g)
    torch.manual_seed(500)
    y4 = d(x4)
    torch.manual_seed(500)
    m4 = torch.empty(6, 10).bernoulli_(1.0 - d.p)  # tensor-p overload, like dropout.py:21-30
    torch.manual_seed(501)
    y2 = d(x2)
    torch.manual_seed(501)
    m2 = torch.empty(16, 50).bernoulli_(1.0 - d.p)
    save("dropout", x4=npy(x4), m4=npy(m4), y4=npy(y4), x2=npy(x2), m2=npy(m2), y2=npy(y2), p=np.array(0.2, np.float32))


def gen_metrics(src):
    import src.metrics as M
    g = torch.Generator().manual_seed(14)
    B, K = 200, 10
    probs = torch.softmax(torch.randn(B, K, generator=g) * 2.0, dim=1)
    target = torch.randint(0, K, (B,), generator=g)
    cm = M.ClassificationMetric(output_size=K)
    cm.update(probs, target)
    # the reference's own ECE binning (experiments/utils.py:293-304 shape): acc/conf per bin
    conf, pred = probs.max(1)
    S, R = 7, 50
    mus = torch.randn(S, R, generator=g)
    vars_ = torch.rand(S, R, generator=g) + 0.1
    tgt = torch.randn(R, generator=g)
    mean = torch.stack(list(mus), dim=1).mean(dim=1)
    var = torch.stack(list(mus), dim=1).var(dim=1) + torch.stack(list(vars_), dim=1).mean(dim=1)
    rm = M.RegressionMetric(output_size=1)
    rm.update((mean, var), tgt)
    plist = torch.softmax(torch.randn(S, 32, K, generator=g), dim=-1)
    pmean = torch.stack(list(plist), dim=1).mean(dim=1)
    save("metrics", probs=npy(probs), target=npy(target), error=npy(cm.error.compute()), nll=npy(cm.nll.compute()),
         brier=npy(cm.brier.compute()), entropy=npy(cm.entropy.compute()), ece=npy(cm.ece.compute()),
         mus=npy(mus), vars=npy(vars_), reg_target=npy(tgt), reg_mean=npy(mean), reg_var=npy(var),
         reg_nll=npy(rm.nll.compute()), reg_mse=npy(rm.mse.compute()), reg_rmse=npy(rm.rmse.compute()),
         reg_mae=npy(rm.mae.compute()), plist=npy(plist), pmean=npy(pmean))


def gen_losses(src):
    """src/losses.py: both tasks, both scalings (losses.py:18-29,37-52), forward values and the gradient w.r.t. the output."""
    import src.losses as L
    g = torch.Generator().manual_seed(15)
    args = Args(loss_multiplier=0.5)
    out = torch.softmax(torch.randn(16, 10, generator=g), dim=1)
    tgt = torch.randint(0, 10, (16,), generator=g)
    loss, ce, kl = L.ClassificationLoss(Args(), "batch")(out, tgt, torch.tensor(123.4), 0.01, 176, 45000)
    arrays = dict(out=npy(out), target=npy(tgt), loss=npy(loss), ce=npy(ce), kl=npy(kl))
    mean, var = torch.randn(16, 3, generator=g), torch.rand(16, 3, generator=g) + 0.05
    rtgt = torch.randn(16, 3, generator=g)
    arrays.update(r_mean=npy(mean), r_var=npy(var), r_target=npy(rtgt))
    for scaling in ("batch", "whole"):
        o = out.clone().requires_grad_(True)
        vals = L.LOSS_FACTORY["classification"](args, scaling)(o, tgt, torch.tensor(123.4), 0.01, 176, 45000)
        vals[0].backward()
        arrays["cls_%s" % scaling], arrays["cls_%s_dout" % scaling] = np.array([v.item() for v in vals]), npy(o.grad)
        m, v = mean.clone().requires_grad_(True), var.clone().requires_grad_(True)
        vals = L.LOSS_FACTORY["regression"](args, scaling)((m, v), rtgt, torch.tensor(55.5), 0.1, 8, 1000)
        vals[0].backward()
        arrays["reg_%s" % scaling] = np.array([x.item() for x in vals])
        arrays["reg_%s_dmean" % scaling], arrays["reg_%s_dvar" % scaling] = npy(m.grad), npy(v.grad)
    save("losses", **arrays)


def _load_resnet_state(net, P):
    """oracle.qbn_oracle.ResNetBBBParams -> reference ConvNetwork_ResNet (names are the reference's)."""
    sd = net.state_dict()
    for name, (mu, rho) in P.convs.items():
        sd[name + ".weight"] = mu
        sd[name + ".std"] = rho
    for name, (w, b, rm, rv, _) in P.bns.items():
        sd[name + ".weight"], sd[name + ".bias"] = w, b
        sd[name + ".running_mean"], sd[name + ".running_var"] = rm, rv
    sd["layers.9.weight"], sd["layers.9.std"] = P.fc
    net.load_state_dict(sd)


def gen_models(src):
    """Whole-network fixtures.  Parameters and noise are NOT stored: they are regenerated from
    seeds by oracle.qbn_oracle.{ResNet,LeNet,MLP}BBBParams / replay_noise (same torch build on the
    GPU box), only the reference outputs are."""
    import oracle.qbn_oracle as O
    from src.models.stochastic.bbb.models_bbb import ConvNetwork_LeNet, ConvNetwork_ResNet, LinearNetwork
    # ---- narrow ResNet-18 (models_bbb.py:191-259), B=4, two replayed eval samples
    P = O.ResNetBBBParams(seed=21)
    args = Args(sigma_prior=0.05, model="conv_resnet_bbb")
    net = ConvNetwork_ResNet([4, 3, 32, 32], 10, False, args)
    _load_resnet_state(net, P)
    x = torch.randn(4, 3, 32, 32, generator=torch.Generator().manual_seed(22))
    net.eval()
    arrays = {}
    for s in range(2):
        torch.manual_seed(700 + s)
        with torch.no_grad():
            arrays["y_eval%d" % s] = npy(net(x))
    arrays["kl"] = npy(net.get_kl_divergence())
    # LRT training forward + backward of the whole net (trainer.py:95-104), loss = ELBO (losses.py:18-29)
    net.train()
    tgt = torch.randint(0, 10, (4,), generator=torch.Generator().manual_seed(23))
    torch.manual_seed(710)
    yt = net(x)
    kl = net.get_kl_divergence()
    loss = F.nll_loss(torch.log(yt + 1e-8), tgt) + 0.01 * kl / (4 * 176)
    loss.backward()
    arrays["y_train"] = npy(yt)
    arrays["loss"] = npy(loss)
    arrays["g.layers.0.weight"] = npy(net.layers[0].weight.grad)
    arrays["g.layers.0.std"] = npy(net.layers[0].std.grad)
    arrays["g.layers.9.weight"] = npy(net.layers[9].weight.grad)
    arrays["g.layers.9.std"] = npy(net.layers[9].std.grad)
    arrays["g.layers.5.0.shortcut.0.weight"] = npy(net.layers[5][0].shortcut[0].weight.grad)
    arrays["g.layers.5.0.shortcut.0.std"] = npy(net.layers[5][0].shortcut[0].std.grad)
    arrays["g.layers.1.weight"] = npy(net.layers[1].weight.grad)
    arrays["bn1.running_mean"] = npy(net.layers[1].running_mean)
    save("resnet", **arrays)

    # ---- LeNet (models_bbb.py:98-143), B=4
    P = O.LeNetBBBParams(seed=31)
    args = Args(sigma_prior=0.1, model="conv_lenet_bbb")
    net = ConvNetwork_LeNet([1, 1, 28, 28], 10, False, args)
    sd = net.state_dict()
    for name, (mu, rho) in P.layers.items():
        sd[name + ".weight"], sd[name + ".std"] = mu, rho
    net.load_state_dict(sd)
    x = torch.rand(4, 1, 28, 28, generator=torch.Generator().manual_seed(32))
    net.eval()
    arrays = {}
    torch.manual_seed(800)
    with torch.no_grad():
        arrays["y_eval0"] = npy(net(x))
    net.train()
    tgt = torch.randint(0, 10, (4,), generator=torch.Generator().manual_seed(33))
    torch.manual_seed(801)
    yt = net(x)
    loss = F.nll_loss(torch.log(yt + 1e-8), tgt) + 0.1 * net.get_kl_divergence() / (4 * 10)
    loss.backward()
    arrays["y_train"] = npy(yt)
    arrays["loss"] = npy(loss)
    arrays["g.layers.0.weight"] = npy(net.layers[0].weight.grad)
    arrays["g.layers.0.std"] = npy(net.layers[0].std.grad)
    arrays["g.layers.7.weight"] = npy(net.layers[7].weight.grad)
    arrays["g.layers.7.std"] = npy(net.layers[7].std.grad)
    save("lenet", **arrays)

    # ---- regression MLP (models_bbb.py:32-96), B=16, config C1
    P = O.MLPBBBParams(seed=41)
    args = Args(sigma_prior=1.0, model="linear_bbb", task="regression")
    net = LinearNetwork([1], 1, False, args)
    sd = net.state_dict()
    for name, (mu, rho, b) in P.layers.items():
        sd[name + ".weight"], sd[name + ".std"], sd[name + ".bias"] = mu, rho, b
    net.load_state_dict(sd)
    x = torch.randn(16, 1, generator=torch.Generator().manual_seed(42))
    net.eval()
    torch.manual_seed(900)
    with torch.no_grad():
        mu, var = net(x)
    save("mlp", y_mu=npy(mu), y_var=npy(var), kl=npy(net.get_kl_divergence()))


def gen_mc_models(src):
    """MC-Dropout networks (configs 2 and 5): reference models_mc.py ConvNetwork_ResNet / ConvNetwork_LeNet, p = 0.15 / 0.2,
    two seeded stochastic forwards each.  Weights: the mu tensors of the seed-generated BBB containers."""
    import oracle.qbn_oracle as O
    from src.models.stochastic.mcdropout.models_mc import ConvNetwork_LeNet, ConvNetwork_ResNet
    P = O.ResNetBBBParams(seed=51)
    args = Args(p=0.15, model="conv_resnet_mc")
    net = ConvNetwork_ResNet([4, 3, 32, 32], 10, False, args)
    sd = net.state_dict()
    sd.update(O.resnet_mc_state_dict(P))
    net.load_state_dict(sd)
    net.eval()
    x = torch.randn(4, 3, 32, 32, generator=torch.Generator().manual_seed(52))
    arrays = {}
    for s in range(2):
        torch.manual_seed(1100 + s)
        with torch.no_grad():
            arrays["y%d" % s] = npy(net(x))
    save("resnet_mc", **arrays)

    P = O.LeNetBBBParams(seed=61)
    args = Args(p=0.2, model="conv_lenet_mc")
    net = ConvNetwork_LeNet([1, 1, 28, 28], 10, False, args)
    sd = net.state_dict()
    for ref_name, key in (("layers.0", "layers.0"), ("layers.3", "layers.2"), ("layers.7", "layers.5"), ("layers.10", "layers.7")):
        sd[ref_name + ".weight"] = P.layers[key][0]
    net.load_state_dict(sd)
    net.eval()
    x = torch.rand(4, 1, 28, 28, generator=torch.Generator().manual_seed(62))
    arrays = {}
    for s in range(2):
        torch.manual_seed(1200 + s)
        with torch.no_grad():
            arrays["y%d" % s] = npy(net(x))
    save("lenet_mc", **arrays)


def _tiny_qnet(src, args):
    """The reference's ConvNetwork_LeNet forward/fuse_model (models_bbb.py:98-143) over the
    reference's own layer classes, with small layer sizes so the int8 fixture stays small."""
    from src.models.stochastic.bbb.conv import Conv2d
    from src.models.stochastic.bbb.linear import Linear
    from src.models.stochastic.bbb.models_bbb import ConvNetwork_LeNet
    from src.utils import Flatten
    net = ConvNetwork_LeNet([1, 1, 28, 28], 10, True, args)
    sp = args.sigma_prior
    net.layers = torch.nn.ModuleList([
        Conv2d(1, 6, (5, 5), stride=1, padding=2, sigma_prior=sp, bias=False, args=args),
        torch.nn.MaxPool2d(2, 2),
        Conv2d(6, 12, (5, 5), stride=1, padding=2, sigma_prior=sp, bias=False, args=args),
        torch.nn.MaxPool2d(2, 2),
        Flatten(),
        Linear(12 * 7 * 7, 32, sigma_prior=sp, bias=False, args=args),
        torch.nn.ReLU(),
        Linear(32, 10, sigma_prior=sp, bias=False, args=args)])
    return net


def gen_qat_int8(src):
    """Config C3 lifecycle on a LeNet-shaped net: prepare_model (quant_utils.py:112-147), one train
    forward + one eval forward (calibrates every observer, SURVEY §5 subtleties), convert
    (quant_utils.py:62-99), then int8 forwards with replayed noise; per-layer integer I/O captured."""
    import src.quant_utils as qu
    g = torch.Generator().manual_seed(17)
    args = Args(sigma_prior=0.1, model="conv_lenet_bbb", q=True, at=True, activation_precision=7, weight_precision=8)
    net = _tiny_qnet(src, args)
    with torch.no_grad():
        for m in net.modules():
            if hasattr(m, "std") and hasattr(m, "weight"):
                trained_like_(m, g)
                m.std.uniform_(-6.0, -3.0, generator=g)
    net.train()
    qu.prepare_model(net, args)
    x = torch.rand(8, 1, 28, 28, generator=g)

    arrays = {"x": npy(x)}
    caps = {}

    def cap(name):
        def hook(mod, inp, out):
            caps[name] = (inp[0].detach().clone(), out.detach().clone())
        return hook

    mods = dict(net.named_modules())
    qat_names = [n for n, m in net.named_modules() if hasattr(m, "weight_fake_quant")]
    hooks = [mods[n].register_forward_hook(cap(n)) for n in qat_names]

    def observer_state(fq):
        o = fq.activation_post_process
        return np.array([float(o.min_val), float(o.max_val)], np.float32)

    def fq_params(fq):
        return np.array([float(fq.scale), float(fq.zero_point), fq.quant_min, fq.quant_max], np.float64)

    arrays["quant_in_fq"] = np.zeros(4)
    # train-mode forward (first observer call initialises min/max)
    torch.manual_seed(1000)
    y = net(x)
    arrays["tr.quant"] = fq_params(net.quant.activation_post_process)
    for n in qat_names:
        m = mods[n]
        arrays["tr.%s.in" % n] = npy(caps[n][0])
        arrays["tr.%s.out" % n] = npy(caps[n][1])
        arrays["tr.%s.wfq" % n] = fq_params(m.weight_fake_quant)
        arrays["tr.%s.sfq" % n] = fq_params(m.std_fake_quant)
        arrays["tr.%s.afq" % n] = fq_params(m.activation_post_process)
        arrays["tr.%s.aobs" % n] = observer_state(m.activation_post_process)
        arrays["p.%s.weight" % n] = npy(m.weight)
        arrays["p.%s.std" % n] = npy(m.std)
    torch.manual_seed(1000)
    for n in qat_names:
        arrays["tr.%s.eps" % n] = npy(torch.empty(caps[n][1].shape).normal_())
    arrays["tr.y"] = npy(y)
    # eval-mode forward (second observer call: EMA update; calibrates add_weight/mul_noise)
    net.eval()
    torch.manual_seed(1001)
    with torch.no_grad():
        ye = net(x)
    for n in qat_names:
        m = mods[n]
        arrays["ev.%s.in" % n] = npy(caps[n][0])
        arrays["ev.%s.out" % n] = npy(caps[n][1])
        arrays["ev.%s.wfq" % n] = fq_params(m.weight_fake_quant)
        arrays["ev.%s.sfq" % n] = fq_params(m.std_fake_quant)
        arrays["ev.%s.afq" % n] = fq_params(m.activation_post_process)
        arrays["ev.%s.aobs" % n] = observer_state(m.activation_post_process)
        arrays["ev.%s.mulfq" % n] = fq_params(m.mul_noise.activation_post_process)
        arrays["ev.%s.addfq" % n] = fq_params(m.add_weight.activation_post_process)
    torch.manual_seed(1001)
    for n in qat_names:
        arrays["ev.%s.eps" % n] = npy(torch.empty(mods[n].weight.shape).normal_())
    arrays["ev.y"] = npy(ye)
    arrays["qat_names"] = np.array(qat_names)
    arrays["qat_relu"] = np.array(["ReLU" in type(mods[n]).__name__ for n in qat_names])
    for h in hooks:
        h.remove()
    save("tiny_qat", **arrays)

    # ---- convert to int8 and capture per-layer integer I/O
    qu.convert(net)
    net.eval()
    arrays = {"x": npy(x)}
    mods = dict(net.named_modules())
    q_names = [n for n, m in net.named_modules() if hasattr(m, "mul_noise") and hasattr(m, "scale")]
    caps = {}
    hooks = [mods[n].register_forward_hook(cap(n)) for n in q_names]
    torch.manual_seed(1002)
    with torch.no_grad():
        yq = net(x)
    for h in hooks:
        h.remove()
    arrays["quant_qp"] = np.array([float(net.quant.scale), int(net.quant.zero_point)], np.float64)
    from src.models.stochastic.bbb.quantized import NOISE_SCALE, NOISE_ZERO_POINT
    torch.manual_seed(1002)
    for n in q_names:
        m = mods[n]
        eps = torch.empty(m.std.shape).normal_()
        arrays["%s.eps" % n] = npy(eps)
        xin, out = caps[n]
        arrays["%s.x_q" % n] = npy(xin.int_repr())
        arrays["%s.x_qp" % n] = np.array([xin.q_scale(), xin.q_zero_point()], np.float64)
        arrays["%s.y_q" % n] = npy(out.int_repr())
        arrays["%s.y_qp" % n] = np.array([out.q_scale(), out.q_zero_point()], np.float64)
        arrays["%s.mu_q" % n] = npy(m.weight.int_repr())
        arrays["%s.mu_qp" % n] = np.array([m.weight.q_scale(), m.weight.q_zero_point()], np.float64)
        arrays["%s.sigma_q" % n] = npy(m.std.int_repr())
        arrays["%s.sigma_qp" % n] = np.array([m.std.q_scale(), m.std.q_zero_point()], np.float64)
        arrays["%s.mul_qp" % n] = np.array([m.mul_noise.scale, m.mul_noise.zero_point], np.float64)
        arrays["%s.add_qp" % n] = np.array([m.add_weight.scale, m.add_weight.zero_point], np.float64)
        arrays["%s.out_qp" % n] = np.array([m.scale, m.zero_point], np.float64)
        arrays["%s.relu" % n] = np.array("ReLU" in type(m).__name__)
        # the sampled int8 weight the module built internally (recomputed with the same ops)
        nq = torch.quantize_per_tensor(eps, NOISE_SCALE, NOISE_ZERO_POINT, dtype=torch.qint8)
        w = m.add_weight.add(m.weight, m.mul_noise.mul(m.std, nq))
        arrays["%s.w_q" % n] = npy(w.int_repr())
        if m.weight.dim() == 4:
            arrays["%s.conv" % n] = np.array([m.stride[0], m.padding[0]])
    arrays["q_names"] = np.array(q_names)
    arrays["y"] = npy(yq)
    save("tiny_int8", **arrays)


def _tiny_qresnet(src, args):
    """The reference's ConvNetwork_ResNet forward/fuse_model (models_bbb.py:191-259) and BasicBlock (:146-188) over the
    reference's own layer classes: stem conv + one identity block + one stride-2 block with a 1x1 shortcut, 8x8 inputs."""
    from src.models.stochastic.bbb.conv import Conv2d
    from src.models.stochastic.bbb.linear import Linear
    from src.models.stochastic.bbb.models_bbb import BasicBlock, ConvNetwork_ResNet
    from src.utils import Flatten
    net = ConvNetwork_ResNet([1, 3, 32, 32], 10, True, args)
    sp = args.sigma_prior
    net.layers = torch.nn.ModuleList([
        Conv2d(3, 8, kernel_size=3, stride=1, padding=1, bias=False, sigma_prior=sp, args=args), torch.nn.BatchNorm2d(8), torch.nn.ReLU(),
        torch.nn.ModuleList([BasicBlock(8, 8, 1, True, args), BasicBlock(8, 16, 2, True, args)]),
        torch.nn.AvgPool2d(4), Flatten(), Linear(16, 10, sigma_prior=sp, bias=False, args=args)])
    return net


def gen_resnet_int8(src):
    """Config C3 on a ResNet-shaped net (conv+BN+ReLU fusion, BN fold at convert, quantised residual add, quantised
    ReLU / average pool): prepare_model, one train + one eval forward, convert, one int8 forward with replayed noise.
    Every int8 module's integer input/output, every residual add and the pooled map are captured in call order."""
    import src.quant_utils as qu
    from src.models.stochastic.bbb.models_bbb import BasicBlock
    from src.models.stochastic.bbb.quantized import NOISE_SCALE, NOISE_ZERO_POINT
    from src.utils import Add
    g = torch.Generator().manual_seed(23)
    args = Args(sigma_prior=0.1, model="conv_resnet_bbb", q=True, at=True, activation_precision=7, weight_precision=8)
    net = _tiny_qresnet(src, args)
    arrays = {}
    with torch.no_grad():
        for n, m in net.named_modules():
            if hasattr(m, "std") and hasattr(m, "weight"):
                trained_like_(m, g)
                m.std.uniform_(-6.0, -3.0, generator=g)
                arrays["p.%s.weight" % n], arrays["p.%s.std" % n] = npy(m.weight), npy(m.std)
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.1, generator=g)
                m.running_var.uniform_(0.5, 1.5, generator=g)
                m.weight.uniform_(0.5, 1.5, generator=g)
                m.bias.normal_(0, 0.1, generator=g)
                for k in ("running_mean", "running_var", "weight", "bias"):
                    arrays["p.%s.%s" % (n, k)] = npy(getattr(m, k))
    net.train()
    qu.prepare_model(net, args)
    for m in net.modules():                  # keep the BatchNorm statistics fixed so the fixture's parameters are the ones used
        if hasattr(m, "freeze_bn_stats"):
            m.freeze_bn_stats()
    x = torch.rand(8, 3, 8, 8, generator=g)
    arrays["x"] = npy(x)
    qat_names = [n for n, m in net.named_modules() if hasattr(m, "weight_fake_quant")]     # = call order (stem, stem, shortcut)
    qat_mods = dict(net.named_modules())
    out_shapes = {}
    hooks = [qat_mods[n].register_forward_hook(lambda m, i, o, n=n: out_shapes.__setitem__(n, tuple(o.shape))) for n in qat_names]
    torch.manual_seed(2000)
    arrays["tr.y"] = npy(net(x))
    for h in hooks:
        h.remove()
    torch.manual_seed(2000)
    for n in qat_names:                      # LRT noise has the layer-output shape, one draw per layer in call order
        arrays["tr.%s.eps" % n] = npy(torch.empty(out_shapes[n]).normal_())
    net.eval()
    torch.manual_seed(2001)
    with torch.no_grad():
        arrays["ev.y"] = npy(net(x))
    torch.manual_seed(2001)
    for n in qat_names:                      # eval draws one weight-shaped eps per layer
        arrays["ev.%s.eps" % n] = npy(torch.empty(qat_mods[n].weight.shape).normal_())
    arrays["qat_names"] = np.array(qat_names)
    arrays["qat_types"] = np.array([type(dict(net.named_modules())[n]).__name__ for n in qat_names])

    qu.convert(net)
    net.eval()
    mods = dict(net.named_modules())
    order, caps = [], {}

    def cap(name):
        def hook(mod, inp, out):
            order.append(name)
            caps[name] = (tuple(t.detach().clone() for t in inp), out.detach().clone())
        return hook

    watched = [n for n, m in mods.items() if (hasattr(m, "mul_noise") and hasattr(m, "scale")) or isinstance(m, (Add, BasicBlock, torch.nn.AvgPool2d))]
    hooks = [mods[n].register_forward_hook(cap(n)) for n in watched]
    torch.manual_seed(2002)
    with torch.no_grad():
        yq = net(x)
    for h in hooks:
        h.remove()

    def qp(t):
        return np.array([t.q_scale(), t.q_zero_point()], np.float64)

    arrays["quant_qp"] = np.array([float(net.quant.scale), int(net.quant.zero_point)], np.float64)
    q_names = [n for n in order if hasattr(mods[n], "mul_noise")]
    torch.manual_seed(2002)
    for n in q_names:                        # noise replay in call order (stem, stem, shortcut inside a block)
        m = mods[n]
        eps = torch.empty(m.std.shape).normal_()
        (xin,), out = caps[n]
        arrays["%s.eps" % n] = npy(eps)
        arrays["%s.x_q" % n], arrays["%s.x_qp" % n] = npy(xin.int_repr()), qp(xin)
        arrays["%s.y_q" % n], arrays["%s.y_qp" % n] = npy(out.int_repr()), qp(out)
        arrays["%s.mu_q" % n], arrays["%s.mu_qp" % n] = npy(m.weight.int_repr()), qp(m.weight)
        arrays["%s.sigma_q" % n], arrays["%s.sigma_qp" % n] = npy(m.std.int_repr()), qp(m.std)
        arrays["%s.mul_qp" % n] = np.array([m.mul_noise.scale, m.mul_noise.zero_point], np.float64)
        arrays["%s.add_qp" % n] = np.array([m.add_weight.scale, m.add_weight.zero_point], np.float64)
        arrays["%s.out_qp" % n] = np.array([m.scale, m.zero_point], np.float64)
        arrays["%s.relu" % n] = np.array("ReLU" in type(m).__name__)
        bias = m.bias() if callable(getattr(m, "bias", None)) else None
        arrays["%s.bias" % n] = npy(bias) if bias is not None else np.zeros(0, np.float32)
        nq = torch.quantize_per_tensor(eps, NOISE_SCALE, NOISE_ZERO_POINT, dtype=torch.qint8)
        arrays["%s.w_q" % n] = npy(m.add_weight.add(m.weight, m.mul_noise.mul(m.std, nq)).int_repr())
        if m.weight.dim() == 4:
            arrays["%s.conv" % n] = np.array([m.stride[0], m.padding[0]])
    for n in order:
        m = mods[n]
        if isinstance(m, Add):
            (a, b), out = caps[n]
            arrays["%s.a_q" % n], arrays["%s.a_qp" % n] = npy(a.int_repr()), qp(a)
            arrays["%s.b_q" % n], arrays["%s.b_qp" % n] = npy(b.int_repr()), qp(b)
            arrays["%s.y_q" % n], arrays["%s.y_qp" % n] = npy(out.int_repr()), qp(out)
        elif isinstance(m, (BasicBlock, torch.nn.AvgPool2d)):
            (xin,), out = caps[n]
            arrays["%s.x_q" % n], arrays["%s.y_q" % n], arrays["%s.y_qp" % n] = npy(xin.int_repr()), npy(out.int_repr()), qp(out)
    arrays["order"] = np.array(order)
    arrays["q_names"] = np.array(q_names)
    arrays["y"] = npy(yq)
    save("tiny_resnet_int8", **arrays)
    # the checkpoint the reference itself writes for this model (postprocess_model -> utils.save_model: torch.save of the
    # converted state-dict, quant_utils.py:101-110): int8 tensors as torch per-tensor-affine qint8, qparams as 0-d tensors
    path = os.path.join(GOLD, "tiny_resnet_int8_weights.pt")
    torch.save(net.state_dict(), path)
    print("wrote %-28s %7.1f KB" % (os.path.basename(path), os.path.getsize(path) / 1024))



def gen_full_resnet_int8(src):
    """Config C5 at FULL size: the reference's own ConvNetwork_ResNet (24/48/96/192, models_bbb.py:191-259) quantised
    A7/W8 — prepare_model, one train + one eval forward (B=16) to calibrate every observer, convert — then TWO int8 forwards
    of a B=4 batch on FBGEMM with replayed noise.  The fixture holds the reference's own checkpoint of the converted model
    (resnet_int8_full_weights.pt: qint8 weights, every qparam), the input, the class probabilities, every BasicBlock output /
    pooled map / logits in full, and a SHA-1 of every int8 layer's integer output (21 layers x 2 forwards).  The noise is
    NOT stored (6.3 MB per forward): the tests redraw it from torch's CPU generator with the recorded seeds, layer by layer
    in call order, exactly as below; `eps_sha1` pins that the generator still yields the same stream."""
    import hashlib
    import src.quant_utils as qu
    from src.models.stochastic.bbb.models_bbb import BasicBlock, ConvNetwork_ResNet
    g = torch.Generator().manual_seed(31)
    args = Args(sigma_prior=0.05, model="conv_resnet_bbb", q=True, at=True, activation_precision=7, weight_precision=8)
    torch.manual_seed(30)
    net = ConvNetwork_ResNet([1, 3, 32, 32], 10, True, args)
    with torch.no_grad():
        for n, m in net.named_modules():
            if hasattr(m, "std") and hasattr(m, "weight"):
                trained_like_(m, g)
                m.std.uniform_(-6.0, -3.0, generator=g)
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.1, generator=g)
                m.running_var.uniform_(0.5, 1.5, generator=g)
                m.weight.uniform_(0.5, 1.5, generator=g)
                m.bias.normal_(0, 0.1, generator=g)
    net.train()
    qu.prepare_model(net, args)
    for m in net.modules():
        if hasattr(m, "freeze_bn_stats"):
            m.freeze_bn_stats()
    torch.set_num_threads(8)
    xc = torch.randn(16, 3, 32, 32, generator=g)
    torch.manual_seed(3000)
    net(xc)
    net.eval()
    torch.manual_seed(3001)
    with torch.no_grad():
        net(xc)
    qu.convert(net)
    net.eval()
    torch.set_num_threads(1)
    mods = dict(net.named_modules())
    x = torch.randn(4, 3, 32, 32, generator=g)
    arrays = {"x": npy(x), "seeds": np.array([3002, 3003])}
    order, caps = [], {}

    def cap(name):
        def hook(mod, inp, out):
            order.append(name)
            caps[name] = out.detach().clone()
        return hook
    watched = [n for n, m in mods.items() if (hasattr(m, "mul_noise") and hasattr(m, "scale")) or isinstance(m, (BasicBlock, torch.nn.AvgPool2d))]
    hooks = [mods[n].register_forward_hook(cap(n)) for n in watched]
    sha = lambda a: hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()  # noqa: E731
    for fi, seed in enumerate((3002, 3003)):
        order.clear()
        caps.clear()
        torch.manual_seed(seed)
        with torch.no_grad():
            y = net(x)
        arrays["f%d.y" % fi] = npy(y)
        q_names = [n for n in order if hasattr(mods[n], "mul_noise")]
        torch.manual_seed(seed)
        eh = hashlib.sha1()
        for n in q_names:
            eh.update(np.ascontiguousarray(npy(torch.empty(mods[n].std.shape).normal_())).tobytes())
        arrays["f%d.eps_sha1" % fi] = np.array(eh.hexdigest())
        for n in order:
            out = caps[n]
            ints = npy(out.int_repr())                        # logical NCHW order
            arrays["f%d.%s.sha1" % (fi, n)] = np.array(sha(ints.astype(np.uint8)))
            # what the NEXT module sees: the model runs clamp_activation on every module output (models_bbb.py:231-238,
            # src/utils.py:25-30: integer clamp to [0, 2^a - 1], qparams unchanged)
            arrays["f%d.%s.sha1_clamped" % (fi, n)] = np.array(sha(np.clip(ints, 0, 127).astype(np.uint8)))
            arrays["f%d.%s.qp" % (fi, n)] = np.array([out.q_scale(), out.q_zero_point()], np.float64)
            if not hasattr(mods[n], "mul_noise") or mods[n].weight.dim() == 2:
                arrays["f%d.%s.y_q" % (fi, n)] = ints.astype(np.uint8)
        if fi == 0:
            arrays["order"] = np.array(order)
            arrays["q_names"] = np.array(q_names)
    for h in hooks:
        h.remove()
    arrays["quant_qp"] = np.array([float(net.quant.scale), int(net.quant.zero_point)], np.float64)
    save("resnet_int8_full", **arrays)
    path = os.path.join(GOLD, "resnet_int8_full_weights.pt")
    torch.save(net.state_dict(), path)
    print("wrote %-28s %7.1f KB" % (os.path.basename(path), os.path.getsize(path) / 1024))


def gen_quant_ops(src):
    """Direct pins of the third-party integer ops (torch.ops.quantized.*) the A6 recipe calls."""
    rng = np.random.default_rng(18)
    arrays = {}
    n = 4096 + 37
    s_mu, z_mu, s_sig, z_sig = 3.1e-3, -7, 2.2e-4, -128
    s_mul, z_mul, s_add, z_add = 6.3e-4, 3, 3.3e-3, -5
    eps = torch.randn(n, generator=torch.Generator().manual_seed(1))
    mu_i = torch.as_tensor(rng.integers(-128, 128, n).astype(np.int8))
    sg_i = torch.as_tensor(rng.integers(-128, 128, n).astype(np.int8))
    qmu = torch._make_per_tensor_quantized_tensor(mu_i, s_mu, z_mu)
    qsg = torch._make_per_tensor_quantized_tensor(sg_i, s_sig, z_sig)
    from src.models.stochastic.bbb.quantized import NOISE_SCALE, NOISE_ZERO_POINT
    qe = torch.quantize_per_tensor(eps, NOISE_SCALE, NOISE_ZERO_POINT, dtype=torch.qint8)
    r = torch.ops.quantized.mul(qsg, qe, s_mul, z_mul)
    w = torch.ops.quantized.add(qmu, r, s_add, z_add)
    arrays.update(eps=npy(eps), mu_i=npy(mu_i), sg_i=npy(sg_i), eps_q=npy(qe.int_repr()), r=npy(r.int_repr()), w=npy(w.int_repr()),
                  qp=np.array([s_mu, z_mu, s_sig, z_sig, s_mul, z_mul, s_add, z_add], np.float64))
    # quantized linear / conv with float bias, relu and non-relu
    torch.backends.quantized.engine = "fbgemm"
    B, K, N = 9, 70, 12
    s_x, z_x, s_w, z_w, s_o, z_o = 0.0173, 3, 0.0041, -2, 0.052, 61
    xq = torch.as_tensor(rng.integers(0, 128, (B, K)).astype(np.uint8))
    wq = torch.as_tensor(rng.integers(-128, 128, (N, K)).astype(np.int8))
    bias = torch.as_tensor(rng.normal(0, 0.3, N).astype(np.float32))
    qx = torch._make_per_tensor_quantized_tensor(xq, s_x, z_x)
    qw = torch._make_per_tensor_quantized_tensor(wq, s_w, z_w)
    for relu in (False, True):
        for b in (None, bias):
            pk = torch.ops.quantized.linear_prepack(qw, b)
            op = torch.ops.quantized.linear_relu if relu else torch.ops.quantized.linear
            y = op(qx, pk, s_o, z_o)
            arrays["lin_y_relu%d_bias%d" % (relu, b is not None)] = npy(y.int_repr())
    arrays.update(lin_x=npy(xq), lin_w=npy(wq), lin_bias=npy(bias), lin_qp=np.array([s_x, z_x, s_w, z_w, s_o, z_o], np.float64))
    Bc, C, H, Nc = 2, 6, 9, 8
    xq = torch.as_tensor(rng.integers(0, 128, (Bc, C, H, H)).astype(np.uint8))
    wq = torch.as_tensor(rng.integers(-128, 128, (Nc, C, 3, 3)).astype(np.int8))
    bias = torch.as_tensor(rng.normal(0, 0.3, Nc).astype(np.float32))
    qx = torch._make_per_tensor_quantized_tensor(xq, s_x, z_x)
    qw = torch._make_per_tensor_quantized_tensor(wq, s_w, z_w)
    for stride in (1, 2):
        for relu in (False, True):
            pk = torch.ops.quantized.conv2d_prepack(qw, bias, [stride, stride], [1, 1], [1, 1], 1)
            op = torch.ops.quantized.conv2d_relu if relu else torch.ops.quantized.conv2d
            y = op(qx, pk, s_o, z_o)
            arrays["conv_y_s%d_relu%d" % (stride, relu)] = npy(y.int_repr())
    arrays.update(conv_x=npy(xq), conv_w=npy(wq), conv_bias=npy(bias))
    # quantized add on quint8 (residual add, src/utils.py:49-55)
    a = torch.as_tensor(rng.integers(0, 128, 3000).astype(np.uint8))
    b = torch.as_tensor(rng.integers(0, 128, 3000).astype(np.uint8))
    qa = torch._make_per_tensor_quantized_tensor(a, 0.031, 5)
    qb = torch._make_per_tensor_quantized_tensor(b, 0.017, 9)
    y = torch.ops.quantized.add(qa, qb, 0.044, 11)
    arrays.update(add_a=npy(a), add_b=npy(b), add_y=npy(y.int_repr()), add_qp=np.array([0.031, 5, 0.017, 9, 0.044, 11], np.float64))
    # fake quantize forward/backward mask
    xf = torch.randn(5000, generator=torch.Generator().manual_seed(2)) * 0.3
    xf.requires_grad_(True)
    yf = torch.fake_quantize_per_tensor_affine(xf, 0.0047, -3, -128, 127)
    yf.backward(torch.ones_like(yf))
    arrays.update(fq_x=npy(xf), fq_y=npy(yf), fq_mask=npy(xf.grad), fq_qp=np.array([0.0047, -3, -128, 127], np.float64))
    # int8 dropout (dropout.py:31-39)
    from src.models.stochastic.mcdropout.dropout import BernoulliDropout
    d = BernoulliDropout(0.15)
    d.mul_mask = torch.nn.quantized.QFunctional()
    d.mul_mask.scale, d.mul_mask.zero_point = 0.023, 0
    d.mul_scalar = torch.nn.quantized.QFunctional()
    xq = torch.as_tensor(rng.integers(0, 128, (4, 6, 5, 5)).astype(np.uint8))
    qx = torch._make_per_tensor_quantized_tensor(xq, 0.019, 2)
    torch.manual_seed(77)
    y = d(qx)
    torch.manual_seed(77)
    mask = torch.empty(4, 6).bernoulli_(1.0 - d.p)
    arrays.update(do_x=npy(xq), do_mask=npy(mask), do_y=npy(y.int_repr()), do_y_qp=np.array([y.q_scale(), y.q_zero_point()], np.float64),
                  do_qp=np.array([0.019, 2, 0.023, 0, float(d.multiplier)], np.float64))
    save("quant_ops", **arrays)


def main():
    src = import_reference()
    torch.set_num_threads(1)  # deterministic vector/tail split in ATen's quantised kernels
    gen_linear(src)
    gen_conv(src)
    gen_dropout(src)
    gen_metrics(src)
    gen_losses(src)
    gen_models(src)
    gen_mc_models(src)
    gen_quant_ops(src)
    gen_qat_int8(src)
    gen_resnet_int8(src)
    gen_full_resnet_int8(src)


if __name__ == "__main__":
    import sys
    if len(sys.argv) > 1:                  # regenerate selected fixtures only, e.g. `python oracle/make_golden.py gen_mc_models`
        _src = import_reference()
        torch.set_num_threads(1)
        for _name in sys.argv[1:]:
            globals()[_name](_src)
    else:
        main()
