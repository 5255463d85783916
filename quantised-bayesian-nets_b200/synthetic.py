"""Seed-generated ("synthetic") parameter containers of the reference's model families, in the reference's own layouts
(OIHW mu/rho, BatchNorm running statistics), in trained-like ranges as SURVEY 8d prescribes.  Shared by the benchmarks
(bench.py, scripts/), the drop-in loaders (zoo.*_from_params) and — re-exported by oracle/qbn_oracle.py — the parity
tests and golden-fixture generator, so every party builds bit-identical parameters from a seed.  Pure torch, no kernels."""
import math

import numpy as np
import torch


class ResNetBBBParams:
    """Random-init parameters of models_bbb.py:191-259 (narrow ResNet-18: 24/48/96/192) in the
    reference's own layout (OIHW mu/rho, BatchNorm running stats).  `trained_like` perturbs them
    to trained-like ranges as SURVEY §8d prescribes so that activations are not degenerate."""

    def __init__(self, in_ch=3, n_classes=10, seed=1, trained_like=True):
        g = torch.Generator().manual_seed(seed)
        self.convs = {}
        self.bns = {}
        self.blocks = []

        def conv(name, cin, cout, k):
            if trained_like:
                mu = torch.randn(cout, cin, k, k, generator=g) * (1.0 / math.sqrt(cin * k * k))
                rho = torch.empty(cout, cin, k, k).uniform_(-6.0, -4.0, generator=g)
            else:
                mu = torch.empty(cout, cin, k, k).uniform_(-0.01, 0.01, generator=g)  # conv.py:15
                rho = torch.full((cout, cin, k, k), -10.0)  # conv.py:16-17
            self.convs[name] = (mu, rho)

        def bn(name, c):
            w = torch.empty(c).uniform_(0.5, 1.5, generator=g)
            b = torch.randn(c, generator=g) * 0.1
            rm = torch.randn(c, generator=g) * 0.1
            rv = torch.empty(c).uniform_(0.5, 1.5, generator=g)
            self.bns[name] = (w, b, rm, rv, 1e-5)

        conv("layers.0", in_ch, 24, 3)
        bn("layers.1", 24)
        in_planes = 24
        for li, (planes, stride0) in enumerate([(24, 1), (48, 2), (96, 2), (192, 2)]):
            for bi, stride in enumerate([stride0, 1]):
                p = "layers.%d.%d" % (3 + li, bi)
                conv(p + ".stem.0", in_planes, planes, 3)
                bn(p + ".stem.1", planes)
                conv(p + ".stem.3", planes, planes, 3)
                bn(p + ".stem.4", planes)
                has_sc = stride != 1 or in_planes != planes
                if has_sc:
                    conv(p + ".shortcut.0", in_planes, planes, 1)
                    bn(p + ".shortcut.1", planes)
                self.blocks.append((p, stride, has_sc))
                in_planes = planes
        if trained_like:
            self.fc = (torch.randn(n_classes, 192, generator=g) * (1.0 / math.sqrt(192)),
                       torch.empty(n_classes, 192).uniform_(-6.0, -4.0, generator=g))
        else:
            self.fc = (torch.empty(n_classes, 192).uniform_(-0.01, 0.01, generator=g), torch.full((n_classes, 192), -3.0))


class LeNetBBBParams:
    """models_bbb.py:98-143 (ConvNetwork_LeNet) parameters in trained-like ranges."""

    def __init__(self, seed=1, in_ch=1, n_classes=10):
        g = torch.Generator().manual_seed(seed)
        self.layers = {}
        for name, shape in (("layers.0", (20, in_ch, 5, 5)), ("layers.2", (50, 20, 5, 5)),
                            ("layers.5", (500, 50 * 7 * 7)), ("layers.7", (n_classes, 500))):
            fan_in = int(np.prod(shape[1:]))
            mu = torch.randn(shape, generator=g) / math.sqrt(fan_in)
            rho = torch.empty(shape).uniform_(-6.0, -4.0, generator=g)
            self.layers[name] = (mu, rho)

    def noise_plan(self):
        return [(k, tuple(v[0].shape)) for k, v in self.layers.items()]


class MLPBBBParams:
    """models_bbb.py:32-96 (LinearNetwork 1-100-100-100-(mu,log_var)), bias=True everywhere."""

    def __init__(self, seed=1, in_features=1):
        g = torch.Generator().manual_seed(seed)
        self.layers = {}
        for name, (n, k) in (("layers.0", (100, in_features)), ("layers.2", (100, 100)), ("layers.4", (100, 100)),
                             ("mu", (1, 100)), ("log_var", (1, 100))):
            mu = torch.randn(n, k, generator=g) / math.sqrt(k)
            rho = torch.empty(n, k).uniform_(-5.0, -3.0, generator=g)
            b = torch.randn(n, generator=g) * 0.1
            self.layers[name] = (mu, rho, b)

    def noise_plan(self):
        return [(k, tuple(v[0].shape)) for k, v in self.layers.items()]


def resnet_mc_state_dict(P):
    """ResNetBBBParams -> state-dict entries under the module names of models_mc.py's ConvNetwork_ResNet
    (one more top-level module — the dropout after layers.0-2 — and dropouts inside the stem shift the indices)."""
    sd = {"layers.0.weight": P.convs["layers.0"][0]}
    w, b, rm, rv, _ = P.bns["layers.1"]
    sd.update({"layers.1.weight": w, "layers.1.bias": b, "layers.1.running_mean": rm, "layers.1.running_var": rv})
    for pfx, _, has_sc in P.blocks:
        li, bi = int(pfx.split(".")[1]), int(pfx.split(".")[2])
        q = "layers.%d.%d" % (li + 1, bi)
        for src, dst in ((".stem.0", ".stem.0"), (".stem.3", ".stem.4"), (".shortcut.0", ".shortcut.0")):
            if pfx + src in P.convs:
                sd[q + dst + ".weight"] = P.convs[pfx + src][0]
        for src, dst in ((".stem.1", ".stem.1"), (".stem.4", ".stem.5"), (".shortcut.1", ".shortcut.1")):
            if pfx + src in P.bns:
                w, b, rm, rv, _ = P.bns[pfx + src]
                sd.update({q + dst + ".weight": w, q + dst + ".bias": b, q + dst + ".running_mean": rm, q + dst + ".running_var": rv})
    sd["layers.10.weight"] = P.fc[0]
    return sd
