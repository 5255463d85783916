"""Drop-in for src/models/stochastic/bbb/conv.py (Conv2d, the fused containers and BN folding)."""
import torch
import torch.nn as nn
from torch.nn import ReLU

from ... import config, noise, ops
from ._shared import attach_bayes_state, check_bn_fusable, fold_batchnorm, folded_copy, noise_key, typed_container
from .linear import eval_forward
from .utils_bbb import kl_divergence_from_rho


class Conv2d(nn.Conv2d):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=False, padding_mode='zeros', sigma_prior=-2, args=None):
        if groups != 1 or padding_mode != 'zeros':
            raise NotImplementedError("qbn_b200 Conv2d: groups=1, zero padding (all the reference models use)")
        nn.Conv2d.__init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, padding_mode)
        # conv.py:15-18: rho = -10 everywhere; the prior is `tensor((1,)) * sigma_prior` (keeps sigma_prior's dtype)
        attach_bayes_state(self, -10.0, torch.tensor((1,)) * sigma_prior, args)

    _key = noise_key

    def forward(self, X):
        if not self.training:
            return eval_forward(self, X, self.stride, self.padding, self.dilation)      # conv.py:33-39
        # conv.py:24-32.  NOTE the reference adds a [N] bias to an NCHW tensor without reshaping (conv.py:32), which only
        # broadcasts when Wo == N; every reference model uses bias=False.  Here the bias is added per output channel.
        # the requested mode: LRTFunction takes the planar tcgen05 path when the shape allows (TF32; C zero-padded to a multiple of 8,
        # so the 3-channel first layer too), else the gather kernels in the mode config.pick_math_mode allows for this shape
        mode = config.math_mode()
        return ops.LRTFunction.apply(X, self.weight, self.std, self.bias, self.stride, self.padding, self.dilation,
                                     noise.pop_injected(), self._key(), mode, False, None)

    def get_kl_divergence(self):
        """conv.py:43-47."""
        return kl_divergence_from_rho(self.weight, self.std, self.std_prior)


ConvBn2d = typed_container("ConvBn2d", "Conv2d + BatchNorm2d awaiting QAT / folding (conv.py:49-54).", Conv2d, nn.BatchNorm2d, module=__name__)
ConvReLU2d = typed_container("ConvReLU2d", "Conv2d + ReLU (conv.py:56-61).", Conv2d, ReLU, module=__name__)
ConvBnReLU2d = typed_container("ConvBnReLU2d", "Conv2d + BatchNorm2d + ReLU (conv.py:63-68).", Conv2d, nn.BatchNorm2d, ReLU, module=__name__)

# public names of the reference's folding helpers (conv.py:70-88; conv_q.py:130 calls the first)
fuse_conv_bn_weights = fold_batchnorm
fuse_conv_bn_eval = folded_copy


def _fuse(conv, bn, relu=None):
    """torch.quantization.fuse_modules hooks (conv.py:90-115): training -> a typed container for prepare_qat to swap,
    eval -> BatchNorm folded into (mu, rho, bias)."""
    parts = (conv, bn) if relu is None else (conv, bn, relu)
    if len({m.training for m in parts}) != 1:
        raise AssertionError("Conv and BN both must be in the same mode (train or eval).")
    if relu is not None and type(conv) is not Conv2d:
        raise NotImplementedError("Cannot fuse %s modules: %s" % ("train" if conv.training else "eval", (conv, bn, relu)))
    if conv.training:
        check_bn_fusable(conv, bn)
        return ConvBn2d(conv, bn) if relu is None else ConvBnReLU2d(conv, bn, relu)
    folded = folded_copy(conv, bn)
    return folded if relu is None else ConvReLU2d(folded, relu)


def fuse_conv_bn(conv, bn):
    return _fuse(conv, bn)


def fuse_conv_bn_relu(conv, bn, relu):
    return _fuse(conv, bn, relu)
