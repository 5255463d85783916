"""Drop-in for src/models/stochastic/bbb/linear.py: same class names, constructor and forward
signatures, parameter names (`weight`=mu, `std`=rho, `std_prior`, `bias`) and FloatFunctional
attributes (`add_weight`, `mul_noise`), so reference checkpoints load both ways."""
import torch
import torch.nn as nn
from torch.nn import ReLU

from ... import config, noise, ops
from ._shared import attach_bayes_state, noise_key, typed_container
from .utils_bbb import kl_divergence


class Linear(nn.Linear):
    def __init__(self, in_features, out_features, bias, sigma_prior=1.0, args=None):
        nn.Linear.__init__(self, in_features, out_features, bias)
        # linear.py:11-19: rho = -3 everywhere, prior std as a float32 [1] tensor, small uniform bias
        attach_bayes_state(self, -3.0, torch.ones((1,)) * sigma_prior, args)
        if self.bias is not None:
            with torch.no_grad():
                self.bias.uniform_(-0.01, 0.01)

    _key = noise_key

    def get_kl_divergence(self):
        """linear.py:24-28."""
        return kl_divergence(self.weight, self.std, None, self.std_prior)

    def forward(self, x):
        vector = x.dim() == 1
        rows = x.unsqueeze(0) if vector else x
        if self.training:
            # linear.py:32-40 — LRT: both contractions + noise + bias in one kernel
            mode = config.pick_math_mode(self.in_features, self.out_features, lrt=True)
            out = ops.LRTFunction.apply(rows, self.weight, self.std, self.bias, 1, 0, 1, noise.pop_injected(), self._key(), mode, False, None)
        else:
            # linear.py:42-50 — one weight draw per forward, W = mu + softplus(rho)*eps
            out = eval_forward(self, rows.detach(), 1, 0, 1)
        return out[0] if vector else out


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


LinearReLU = typed_container("LinearReLU", "Linear + ReLU awaiting QAT / conversion (linear.py:54-59).", Linear, ReLU, module=__name__)
