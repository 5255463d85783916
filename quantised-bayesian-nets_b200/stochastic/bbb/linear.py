"""Drop-in for src/models/stochastic/bbb/linear.py: same class names, constructor and forward
signatures, parameter names (`weight`=mu, `std`=rho, `std_prior`, `bias`) and FloatFunctional
attributes (`add_weight`, `mul_noise`), so reference checkpoints load both ways."""
import torch
import torch.nn as nn
from torch.nn import ReLU

from ... import config, noise, ops
from .utils_bbb import kl_divergence


class Linear(nn.Linear):
    def __init__(self, in_features, out_features, bias, sigma_prior=1.0, args=None):
        super(Linear, self).__init__(in_features, out_features, bias)
        self.std_prior = torch.nn.Parameter(torch.ones((1,)) * sigma_prior, requires_grad=False)  # linear.py:11-12
        self.weight.data.uniform_(-0.01, 0.01)                                                    # linear.py:14
        self.std = nn.Parameter(torch.zeros_like(self.weight).uniform_(-3, -3))                   # linear.py:15
        self.add_weight = torch.ao.nn.quantized.FloatFunctional()
        self.mul_noise = torch.ao.nn.quantized.FloatFunctional()
        self.args = args
        if self.bias is not None:
            self.bias.data.uniform_(-0.01, 0.01)
        self._qbn_layer_id = noise.new_layer_id()

    def get_kl_divergence(self):
        """linear.py:24-28."""
        return kl_divergence(self.weight, self.std, None, self.std_prior)

    def _key(self):
        return (noise.seed(), self._qbn_layer_id, noise.next_draw())

    def forward(self, x):
        squeeze = x.dim() == 1
        if squeeze:
            x = x.unsqueeze(0)
        K, N = self.in_features, self.out_features
        if self.training:
            # linear.py:32-40 — LRT: both contractions + noise + bias in one kernel
            mode = config.pick_math_mode(K, N, lrt=True)
            eps = noise.pop_injected()
            out = ops.LRTFunction.apply(x, self.weight, self.std, self.bias, 1, 0, 1, eps, self._key(), mode, False, None)
        else:
            # linear.py:42-50 — one weight draw per forward, W = mu + softplus(rho)*eps
            out = eval_forward(self, x.detach(), 1, 0, 1)
        return out.squeeze(0) if squeeze else out


def eval_forward(mod, x, stride, padding, dilation, relu=False):
    """Shared by Linear/Conv2d: sample W (A4) then contract; bias rides the epilogue."""
    with torch.no_grad():
        xc = ops.nhwc(ops._f32(x))
        d = ops._geom(xc, mod.weight.shape, stride, padding, dilation)
        packed = ops.weight_prep(mod.weight, mod.std, False, None, want=("mu", "sigma"))
        eps = noise.pop_injected()
        eps_p = ops.pack_ohwi(ops._f32(eps)).reshape(1, -1) if eps is not None else None
        seed, lid, draw = mod._key()
        mode = config.pick_math_mode(d.C, d.N, lrt=False)
        w = ops.sample_weights(packed["mu"], packed["sigma"], 1, eps_p, seed, lid, draw, round_tf32=(mode == ops.QBN_MATH_TF32))
        return ops.conv_forward(xc, w, d, 1, True, False, None, mod.bias, None, relu, None, 1.0, mode)


class LinearReLU(torch.nn.Sequential):
    def __init__(self, linear, relu):
        assert type(linear) == Linear and type(relu) == ReLU, \
            'Incorrect types for input modules{}{}'.format(type(linear), type(relu))
        super(LinearReLU, self).__init__(linear, relu)
