"""Host-side mirror of the reference's model containers that sit directly on top of the stochastic
layers (src/models/stochastic/bbb/models_bbb.py, .../mcdropout/models_mc.py, src/utils.py:25-55).

Only the composition is mirrored (same attribute names -> identical state-dict keys, same forward
order -> identical noise order); all arithmetic is in the qbn_b200 layers.  These classes exist so
that the hot path can be run, tested and benchmarked on a box where /root/reference is absent."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import noise, ops
from .quant_utils import QTensor
from .stochastic.bbb.conv import Conv2d, ConvReLU2d, fuse_conv_bn, fuse_conv_bn_relu  # noqa: F401
from .stochastic.bbb.linear import Linear, LinearReLU  # noqa: F401
from .stochastic.bbb.utils_bbb import model_kl_divergence
from .stochastic.mcdropout.dropout import BernoulliDropout

UINT_BOUNDS = {8: [0, 255], 7: [0, 127], 6: [0, 63], 5: [0, 31], 4: [0, 15], 3: [0, 7], 2: [0, 3]}      # src/utils.py:18
INT_BOUNDS = {8: [-128, 127], 7: [-64, 63], 6: [-32, 31], 5: [-16, 15], 4: [-8, 7], 3: [-4, 3], 2: [-2, 1]}  # src/utils.py:19-20


class Args:
    """Stand-in for the argparse Namespace the reference threads through every layer."""

    def __init__(self, **kw):
        self.sigma_prior = 1.0
        self.activation_precision = 7
        self.weight_precision = 8
        self.p = 0.2
        self.q = False
        self.at = False
        self.model = "conv_resnet_bbb"
        self.task = "classification"
        self.samples = 100
        self.__dict__.update(kw)


def clamp_activation(x, args):
    """src/utils.py:25-30: only acts on quantised tensors; float tensors pass through."""
    if hasattr(x, "clamp_activation"):
        return x.clamp_activation(args)
    return x


def _apply(layer, x, args=None):
    """Run a non-stochastic glue layer; on int8 activations (QTensor) pooling/ReLU act on the integers
    (order-preserving per-tensor affine map), like torch's quantised max_pool2d / relu / avg_pool2d."""
    if not isinstance(x, QTensor):
        return layer(x)
    bits = getattr(args, "activation_precision", 8) if args is not None else 8
    if isinstance(layer, nn.MaxPool2d):
        q = F.max_pool2d(x.q.float(), layer.kernel_size, layer.stride).to(torch.uint8).contiguous(memory_format=torch.channels_last)
        return QTensor(q, x.scale, x.zero_point, x.bits)
    if isinstance(layer, nn.ReLU):
        return QTensor(ops.i8_relu(x.q, x.zero_point, act_bits=bits), x.scale, x.zero_point, min(bits, x.bits))
    if isinstance(layer, nn.AvgPool2d):
        k = layer.kernel_size if isinstance(layer.kernel_size, int) else layer.kernel_size[0]
        return QTensor(ops.i8_avgpool(x.q, x.zero_point, k, act_bits=bits), x.scale, x.zero_point, min(bits, x.bits))
    return layer(x)


class Flatten(nn.Module):
    def forward(self, x):
        if len(x.shape) == 1:
            return x.unsqueeze(dim=0)
        return x.reshape(x.size(0), -1)


class Add(nn.Module):
    def __init__(self):
        super().__init__()
        self.add = torch.ao.nn.quantized.FloatFunctional()

    def forward(self, x, y):
        return self.add.add(x, y)


class LinearNetwork(nn.Module):
    """models_bbb.py:32-96."""

    def __init__(self, input_size, output_size, q, args):
        super().__init__()
        self.args = args
        self.input_size = 1
        for i in input_size:
            self.input_size *= int(i)
        self.output_size = int(output_size)
        widths = [100, 100, 100]
        self.layers = nn.ModuleList([])
        for i in range(len(widths)):
            fan_in = self.input_size if i == 0 else widths[i - 1]
            self.layers.append(Linear(fan_in, widths[i], sigma_prior=args.sigma_prior, bias=True, args=args))
            self.layers.append(nn.ReLU())
        self.mu = Linear(widths[-1], 1, sigma_prior=args.sigma_prior, bias=True, args=args)
        self.log_var = Linear(widths[-1], 1, sigma_prior=args.sigma_prior, bias=True, args=args)
        self.q = q

    def forward(self, x):
        for layer in self.layers:
            x = layer(x)
        return (self.mu(x), self.log_var(x).exp())

    def get_kl_divergence(self):
        return model_kl_divergence(self)


class ConvNetwork_LeNet(nn.Module):
    """models_bbb.py:98-143."""

    def __init__(self, input_size, output_size, q, args):
        super().__init__()
        self.args = args
        c0 = input_size[0]
        sp = args.sigma_prior
        self.layers = nn.ModuleList([
            Conv2d(c0, 20, (5, 5), stride=1, padding=2, sigma_prior=sp, bias=False, args=args),
            nn.MaxPool2d(kernel_size=2, stride=2),
            Conv2d(20, 50, (5, 5), stride=1, padding=2, sigma_prior=sp, bias=False, args=args),
            nn.MaxPool2d(kernel_size=2, stride=2),
            Flatten(),
            Linear(50 * 7 * 7, 500, sigma_prior=sp, bias=False, args=args),
            nn.ReLU(),
            Linear(500, output_size, sigma_prior=sp, bias=False, args=args)])
        self.q = q
        if self.q:
            self.quant = torch.ao.quantization.QuantStub()
            self.dequant = torch.ao.quantization.DeQuantStub()

    def forward(self, x):
        if self.q:
            x = clamp_activation(self.quant(x), self.args)
        for layer in self.layers:
            x = clamp_activation(_apply(layer, x, self.args), self.args)
        if self.q:
            x = self.dequant(x)
        return F.softmax(x, dim=-1)

    def get_kl_divergence(self):
        return model_kl_divergence(self)

    def fuse_model(self):
        """models_bbb.py:142-143: fuse layers 5,6 (Linear + ReLU) into a LinearReLU container."""
        self.layers[5] = LinearReLU(self.layers[5], self.layers[6])
        self.layers[6] = nn.Identity()


class BasicBlock(nn.Module):
    """models_bbb.py:146-188."""
    expansion = 1

    def __init__(self, in_planes, planes, stride=1, q=False, args=None):
        super().__init__()
        self.args = args
        sp = args.sigma_prior
        self.stem = nn.ModuleList([
            Conv2d(in_planes, planes, kernel_size=3, stride=stride, padding=1, bias=False, sigma_prior=sp, args=args),
            nn.BatchNorm2d(planes), nn.ReLU(),
            Conv2d(planes, planes, kernel_size=3, stride=1, padding=1, bias=False, sigma_prior=sp, args=args),
            nn.BatchNorm2d(planes)])
        self.shortcut = nn.ModuleList([])
        if stride != 1 or in_planes != self.expansion * planes:
            self.shortcut.append(Conv2d(in_planes, self.expansion * planes, kernel_size=1, stride=stride, bias=False, sigma_prior=sp, args=args))
            self.shortcut.append(nn.BatchNorm2d(self.expansion * planes))
        self.add = Add()
        self.end = nn.ReLU()

    def forward(self, x):
        out = x
        for layer in self.stem:
            out = clamp_activation(_apply(layer, out, self.args), self.args)
        shortcut = x
        for layer in self.shortcut:
            shortcut = clamp_activation(_apply(layer, shortcut, self.args), self.args)
        if isinstance(out, QTensor) and noise.sample_batch_state() is not None and hasattr(self.add.add, "add_relu"):
            # int8 MC engine: residual add, ReLU and the activation clamp in ONE pass (quantized::add_relu floors at the output
            # zero point, which is what add -> clamp -> ReLU -> clamp computes)
            return clamp_activation(self.add.add.add_relu(out, shortcut), self.args)
        out = clamp_activation(self.add(out, shortcut), self.args)
        return clamp_activation(_apply(self.end, out, self.args), self.args)

    def fuse_model(self):
        """models_bbb.py:182-188: conv+BN+ReLU, conv+BN and the shortcut's conv+BN become one module each (a typed
        container for QAT in training mode, a BN-folded conv in eval mode); the absorbed slots turn into Identity."""
        st = self.stem
        st[0], st[1], st[2] = fuse_conv_bn_relu(st[0], st[1], st[2]), nn.Identity(), nn.Identity()
        st[3], st[4] = fuse_conv_bn(st[3], st[4]), nn.Identity()
        if len(self.shortcut) == 2:
            self.shortcut[0], self.shortcut[1] = fuse_conv_bn(self.shortcut[0], self.shortcut[1]), nn.Identity()


class ConvNetwork_ResNet(nn.Module):
    """models_bbb.py:191-259 — the narrow ResNet-18 (24/48/96/192)."""

    def __init__(self, input_size, output_size, q, args):
        super().__init__()
        self.args = args
        self.in_planes = 24
        sp = args.sigma_prior
        self.layers = nn.ModuleList([])
        self.layers.append(Conv2d(input_size[1], 24, kernel_size=3, stride=1, padding=1, bias=False, sigma_prior=sp, args=args))
        self.layers.append(nn.BatchNorm2d(24))
        self.layers.append(nn.ReLU())
        for planes, stride in ((24, 1), (48, 2), (96, 2), (192, 2)):
            blocks = []
            for s in (stride, 1):
                blocks.append(BasicBlock(self.in_planes, planes, s, q, args))
                self.in_planes = planes
            self.layers.append(nn.ModuleList(blocks))
        self.layers.append(nn.AvgPool2d(4))
        self.layers.append(Flatten())
        self.layers.append(Linear(192, output_size, sigma_prior=sp, bias=False, args=args))
        self.q = q
        if self.q:
            self.quant = torch.ao.quantization.QuantStub()
            self.dequant = torch.ao.quantization.DeQuantStub()

    def forward(self, x):
        if self.q:
            x = clamp_activation(self.quant(x), self.args)
        for layer in self.layers:
            for sub in (layer if isinstance(layer, nn.ModuleList) else (layer,)):
                x = clamp_activation(_apply(sub, x, self.args), self.args)
        if self.q:
            x = self.dequant(x)
        return F.softmax(x, dim=-1)

    def fuse_model(self):
        """models_bbb.py:245-249: the input conv+BN+ReLU, then every block."""
        L = self.layers
        L[0], L[1], L[2] = fuse_conv_bn_relu(L[0], L[1], L[2]), nn.Identity(), nn.Identity()
        for m in self.modules():
            if isinstance(m, BasicBlock):
                m.fuse_model()

    def get_kl_divergence(self):
        return model_kl_divergence(self)


# ---- MC-Dropout variants (models_mc.py): stock conv/linear parameters + BernoulliDropout ----------
class BasicBlockMC(nn.Module):
    """models_mc.py:110-157."""

    def __init__(self, in_planes, planes, stride, args):
        super().__init__()
        p = args.p
        self.stem = nn.ModuleList([
            nn.Conv2d(in_planes, planes, 3, stride, 1, bias=False), nn.BatchNorm2d(planes), nn.ReLU(), BernoulliDropout(p),
            nn.Conv2d(planes, planes, 3, 1, 1, bias=False), nn.BatchNorm2d(planes), BernoulliDropout(p)])
        self.shortcut = nn.ModuleList([])
        if stride != 1 or in_planes != planes:
            self.shortcut.extend([nn.Conv2d(in_planes, planes, 1, stride, bias=False), nn.BatchNorm2d(planes), BernoulliDropout(p)])
        self.add = Add()
        self.end = nn.ReLU()

    def forward(self, x):
        out = x
        for layer in self.stem:
            out = layer(out)
        sc = x
        for layer in self.shortcut:
            sc = layer(sc)
        return self.end(self.add(out, sc))


class ConvNetwork_ResNet_MC(nn.Module):
    """models_mc.py:159-226."""

    def __init__(self, input_size, output_size, q, args):
        super().__init__()
        self.args = args
        self.layers = nn.ModuleList([nn.Conv2d(input_size[1], 24, 3, 1, 1, bias=False), nn.BatchNorm2d(24), nn.ReLU(), BernoulliDropout(args.p)])
        in_planes = 24
        for planes, stride in ((24, 1), (48, 2), (96, 2), (192, 2)):
            blocks = []
            for s in (stride, 1):
                blocks.append(BasicBlockMC(in_planes, planes, s, args))
                in_planes = planes
            self.layers.append(nn.ModuleList(blocks))
        self.layers.extend([nn.AvgPool2d(4), Flatten(), nn.Linear(192, output_size, bias=False)])

    def forward(self, x):
        for layer in self.layers:
            if isinstance(layer, nn.ModuleList):
                for sub in layer:
                    x = sub(x)
            else:
                x = layer(x)
        return F.softmax(x, dim=-1)


class ConvNetwork_LeNet_MC(nn.Module):
    """models_mc.py:75-110: conv5x5-dropout-pool, conv5x5-dropout-pool, fc-relu-dropout-fc (config 2)."""

    def __init__(self, input_size, output_size, q, args):
        super().__init__()
        self.args = args
        self.layers = nn.ModuleList([
            nn.Conv2d(input_size[0], 20, 5, padding=2, bias=False), BernoulliDropout(args.p), nn.MaxPool2d(2, 2),
            nn.Conv2d(20, 50, 5, padding=2, bias=False), BernoulliDropout(args.p), nn.MaxPool2d(2, 2),
            Flatten(), nn.Linear(50 * 7 * 7, 500, bias=False), nn.ReLU(), BernoulliDropout(args.p), nn.Linear(500, output_size, bias=False)])

    def forward(self, x):
        for layer in self.layers:
            x = layer(x)
        return F.softmax(x, dim=-1)


# ---- loaders from seed-generated parameter containers (duck-typed: .convs/.bns/.fc/.blocks etc.) ----
def resnet_from_params(P, n_classes=10, args=None):
    args = args or Args(sigma_prior=0.05, model="conv_resnet_bbb")
    net = ConvNetwork_ResNet([1, P.convs["layers.0"][0].shape[1], 32, 32], n_classes, False, args)
    sd = net.state_dict()
    for name, (mu, rho) in P.convs.items():
        sd[name + ".weight"], sd[name + ".std"] = mu, rho
    for name, (w, b, rm, rv, _) in P.bns.items():
        sd[name + ".weight"], sd[name + ".bias"], sd[name + ".running_mean"], sd[name + ".running_var"] = w, b, rm, rv
    sd["layers.9.weight"], sd["layers.9.std"] = P.fc
    net.load_state_dict(sd)
    return net


def lenet_from_params(P, n_classes=10, args=None):
    args = args or Args(sigma_prior=0.1, model="conv_lenet_bbb")
    net = ConvNetwork_LeNet([P.layers["layers.0"][0].shape[1], 1, 28, 28], n_classes, False, args)
    sd = net.state_dict()
    for name, (mu, rho) in P.layers.items():
        sd[name + ".weight"], sd[name + ".std"] = mu, rho
    net.load_state_dict(sd)
    return net


def mlp_from_params(P, args=None):
    args = args or Args(sigma_prior=1.0, model="linear_bbb", task="regression")
    net = LinearNetwork([P.layers["layers.0"][0].shape[1]], 1, False, args)
    sd = net.state_dict()
    for name, (mu, rho, b) in P.layers.items():
        sd[name + ".weight"], sd[name + ".std"], sd[name + ".bias"] = mu, rho, b
    net.load_state_dict(sd)
    return net


def resnet_mc_from_params(P, p=0.15, n_classes=10, state_dict=None):
    """MC-Dropout ResNet (models_mc.py) with the mu tensors of a ResNetBBBParams container as its weights;
    `state_dict`: entries under the reference's module names (default: synthetic.resnet_mc_state_dict(P))."""
    if state_dict is None:
        from .synthetic import resnet_mc_state_dict
        state_dict = resnet_mc_state_dict(P)
    args = Args(p=p, model="conv_resnet_mc")
    net = ConvNetwork_ResNet_MC([1, P.convs["layers.0"][0].shape[1], 32, 32], n_classes, False, args)
    sd = net.state_dict()
    sd.update(state_dict)
    net.load_state_dict(sd)
    return net


def lenet_mc_from_params(P, p=0.2, n_classes=10):
    args = Args(p=p, model="conv_lenet_mc")
    net = ConvNetwork_LeNet_MC([P.layers["layers.0"][0].shape[1], 1, 28, 28], n_classes, False, args)
    sd = net.state_dict()
    for ref_name, key in (("layers.0", "layers.0"), ("layers.3", "layers.2"), ("layers.7", "layers.5"), ("layers.10", "layers.7")):
        sd[ref_name + ".weight"] = P.layers[key][0]
    net.load_state_dict(sd)
    return net
