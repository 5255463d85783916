"""Helpers shared by the Bayesian drop-in modules (float, QAT): parameter set-up, exact-type fused containers, BatchNorm folding,
the QAT plumbing.  They exist so that the per-class files only state what differs between the classes."""
import copy

import torch
import torch.nn.functional as F

from ... import noise


def attach_bayes_state(mod, rho_init, prior, args):
    """The Bayesian part of a layer on top of its nn.Linear / nn.Conv2d base (linear.py:11-19, conv.py:15-21): mu in
    `weight` ~ U(-0.01, 0.01), rho in `std` (constant init), the non-trainable prior std, the two FloatFunctionals whose
    observers quantise the sampled weight, and this layer's Philox stream id."""
    with torch.no_grad():
        mod.weight.uniform_(-0.01, 0.01)
        # uniform_(a, a) rather than fill_: it advances the generator exactly as the reference's init does, so a seeded
        # construction yields the same parameters in both code bases
        rho = torch.empty_like(mod.weight).uniform_(rho_init, rho_init)
    mod.std = torch.nn.Parameter(rho)
    mod.std_prior = torch.nn.Parameter(prior, requires_grad=False)
    for name in ("add_weight", "mul_noise"):
        setattr(mod, name, torch.ao.nn.quantized.FloatFunctional())
    mod.args = args
    mod._qbn_layer_id = noise.new_layer_id()


def noise_key(mod):
    """(seed, layer id, draw counter): one Philox stream per layer and forward."""
    return (noise.seed(), mod._qbn_layer_id, noise.next_draw())


def typed_container(name, doc, *kinds, module=None):
    """nn.Sequential subclass that accepts exactly the given module types, in order (the reference's intrinsic containers
    assert exact types: linear.py:54-59, conv.py:49-68; quantisation mappings key on these classes).  `module` is the
    defining module's __name__, so that instances pickle by reference."""

    def __init__(self, *mods):
        if len(mods) != len(kinds) or any(type(m) is not k for m, k in zip(mods, kinds)):
            raise AssertionError("Incorrect types for input modules" + "".join(str(type(m)) for m in mods))
        torch.nn.Sequential.__init__(self, *mods)

    ns = {"__init__": __init__, "__doc__": doc}
    if module is not None:
        ns["__module__"] = module
    return type(name, (torch.nn.Sequential,), ns)


def fold_batchnorm(mu, bias, rho, mean, var, eps, gamma, beta):
    """BatchNorm folded into a Bayesian conv (conv.py:70-80): with c = gamma / sqrt(var + eps) per output channel,
    mu' = mu * c, sigma' = sigma * c (stored back as rho' = softplus^-1(sigma')), bias' = (bias - mean) * c + beta."""
    from .utils_bbb import softplusinv
    inv = torch.rsqrt(var + eps)
    per_out = (gamma * inv).reshape((-1,) + (1,) * (mu.dim() - 1))
    base = bias if bias is not None else torch.zeros_like(mean)
    new_rho = softplusinv(F.softplus(rho) * per_out)
    pack = [mu * per_out, (base - mean) * inv * gamma + beta, new_rho]
    return tuple(torch.nn.Parameter(t) for t in pack)


def folded_copy(conv, bn):
    """Eval-mode fusion: a deep copy of the conv carrying the BatchNorm (conv.py:82-88)."""
    if conv.training or bn.training:
        raise AssertionError("Fusion only for eval!")
    out = copy.deepcopy(conv)
    out.weight, out.bias, out.std = fold_batchnorm(out.weight, out.bias, out.std, bn.running_mean, bn.running_var, bn.eps, bn.weight, bn.bias)
    return out


def check_bn_fusable(conv, bn):
    problems = []
    if bn.num_features != conv.out_channels:
        problems.append("Output channel of Conv2d must match num_features of BatchNorm2d")
    if not bn.affine:
        problems.append("Only support fusing BatchNorm2d with affine set to True")
    if not bn.track_running_stats:
        problems.append("Only support fusing BatchNorm2d with tracking_running_stats set to True")
    if problems:
        raise AssertionError("; ".join(problems))


class QATMixin:
    """What every QAT variant shares (linear_qat.py:8-70, conv_qat.py:12-80): three fake-quantisers from the qconfig, forward =
    activation observer([ReLU](_forward)), the display name, and `from_float` plumbing that re-uses the float module's
    parameters, FloatFunctionals and Philox id."""
    _FLOAT_MODULE = None
    _RELU = False
    _NAME = None

    def _attach_qat(self, qconfig):
        if not qconfig:
            raise AssertionError("qconfig must be provided for QAT module")
        self.qconfig = qconfig
        self.weight_fake_quant, self.std_fake_quant = qconfig.weight(), qconfig.weight()
        self.activation_post_process = qconfig.activation()

    def _fake_quantised(self, scale=None):
        """(mu~, sigma~): fake-quantised mean and softplus(rho), optionally scaled per output channel first (BN folding)."""
        mu, sigma = self.weight, F.softplus(self.std)
        if scale is not None:
            mu, sigma = mu * scale, sigma * scale
        return self.weight_fake_quant(mu), self.std_fake_quant(sigma)

    def forward(self, x):
        y = self._forward(x)
        return self.activation_post_process(F.relu(y) if self._RELU else y)

    def _get_name(self):
        return self._NAME

    @classmethod
    def _check_float(cls, mod, qconfig):
        if type(mod) != cls._FLOAT_MODULE:
            raise AssertionError("qat." + cls.__name__ + ".from_float only works for " + cls._FLOAT_MODULE.__name__)
        if not qconfig and not getattr(mod, "qconfig", None):
            raise AssertionError("Input float module must have a valid qconfig")

    @staticmethod
    def _adopt(q, src, qconfig):
        """Share the float layer's tensors, observers and Philox id with the new QAT module (no copies); the two
        FloatFunctionals get fresh weight fake-quantisers (linear_qat.py:60-68, conv_qat.py:66-78)."""
        q.activation_post_process = src.activation_post_process
        for attr in ("weight", "std", "std_prior", "bias", "args", "add_weight", "mul_noise", "_qbn_layer_id"):
            setattr(q, attr, getattr(src, attr))
        q.add_weight.activation_post_process = qconfig.weight()
        q.mul_noise.activation_post_process = qconfig.weight()
        return q
