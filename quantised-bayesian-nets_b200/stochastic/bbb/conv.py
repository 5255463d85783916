"""Drop-in for src/models/stochastic/bbb/conv.py (Conv2d, the fused containers and BN folding)."""
import copy

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn import ReLU

from ... import config, noise, ops
from .linear import eval_forward
from .utils_bbb import kl_divergence, softplusinv


class Conv2d(nn.Conv2d):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=False, padding_mode='zeros', sigma_prior=-2, args=None):
        super(Conv2d, self).__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, padding_mode)
        if groups != 1 or padding_mode != 'zeros':
            raise NotImplementedError("qbn_b200 Conv2d: groups=1, zero padding (all the reference models use)")
        self.weight.data.uniform_(-0.01, 0.01)                                                        # conv.py:15
        self.std = torch.nn.Parameter(torch.zeros_like(self.weight).uniform_(-10, -10), requires_grad=True)  # conv.py:16-17
        self.std_prior = torch.nn.Parameter(torch.tensor((1,)) * sigma_prior, requires_grad=False)    # conv.py:18
        self.add_weight = torch.ao.nn.quantized.FloatFunctional()
        self.mul_noise = torch.ao.nn.quantized.FloatFunctional()
        self.args = args
        self._qbn_layer_id = noise.new_layer_id()

    def _key(self):
        return (noise.seed(), self._qbn_layer_id, noise.next_draw())

    def forward(self, X):
        if self.training:
            # conv.py:24-32.  NOTE the reference adds a [N] bias to an NCHW tensor without reshaping
            # (conv.py:32), which only broadcasts when Wo == N; every reference model uses
            # bias=False.  Here the bias is added per output channel.
            mode = config.pick_math_mode(self.in_channels, self.out_channels, lrt=True)
            eps = noise.pop_injected()
            return ops.LRTFunction.apply(X, self.weight, self.std, self.bias, self.stride, self.padding, self.dilation,
                                         eps, self._key(), mode, False, None)
        # conv.py:33-39
        return eval_forward(self, X.detach(), self.stride, self.padding, self.dilation)

    def get_kl_divergence(self):
        """conv.py:43-47."""
        return kl_divergence(self.weight, self.std, None, self.std_prior)


class ConvBn2d(torch.nn.Sequential):
    def __init__(self, conv, bn):
        assert type(conv) == Conv2d and type(bn) == torch.nn.BatchNorm2d, \
            'Incorrect types for input modules{}{}'.format(type(conv), type(bn))
        super(ConvBn2d, self).__init__(conv, bn)


class ConvReLU2d(torch.nn.Sequential):
    def __init__(self, conv, relu):
        assert type(conv) == Conv2d and type(relu) == ReLU, \
            'Incorrect types for input modules{}{}'.format(type(conv), type(relu))
        super(ConvReLU2d, self).__init__(conv, relu)


class ConvBnReLU2d(torch.nn.Sequential):
    def __init__(self, conv, bn, relu):
        assert type(conv) == Conv2d and type(bn) == torch.nn.BatchNorm2d and type(relu) == ReLU, \
            'Incorrect types for input modules{}{}{}'.format(type(conv), type(bn), type(relu))
        super(ConvBnReLU2d, self).__init__(conv, bn, relu)


def fuse_conv_bn_weights(conv_w, conv_b, conv_std, bn_rm, bn_rv, bn_eps, bn_w, bn_b):
    """conv.py:70-80: fold BN into mu AND sigma (sigma through softplusinv(softplus(rho)*c)).
    Weight-sized, runs once at fuse/convert time (not on the per-step path)."""
    if conv_b is None:
        conv_b = bn_rm.new_zeros(bn_rm.shape)
    bn_var_rsqrt = torch.rsqrt(bn_rv + bn_eps)
    c = (bn_w * bn_var_rsqrt).reshape([-1] + [1] * (len(conv_w.shape) - 1))
    conv_w = conv_w * c
    conv_std = softplusinv(F.softplus(conv_std) * c)
    conv_b = (conv_b - bn_rm) * bn_var_rsqrt * bn_w + bn_b
    return torch.nn.Parameter(conv_w), torch.nn.Parameter(conv_b), torch.nn.Parameter(conv_std)


def fuse_conv_bn_eval(conv, bn):
    assert (not (conv.training or bn.training)), "Fusion only for eval!"
    fused_conv = copy.deepcopy(conv)
    fused_conv.weight, fused_conv.bias, fused_conv.std = fuse_conv_bn_weights(
        fused_conv.weight, fused_conv.bias, fused_conv.std, bn.running_mean, bn.running_var, bn.eps, bn.weight, bn.bias)
    return fused_conv


def fuse_conv_bn(conv, bn):
    assert (conv.training == bn.training), "Conv and BN both must be in the same mode (train or eval)."
    if conv.training:
        assert bn.num_features == conv.out_channels, 'Output channel of Conv2d must match num_features of BatchNorm2d'
        assert bn.affine, 'Only support fusing BatchNorm2d with affine set to True'
        assert bn.track_running_stats, 'Only support fusing BatchNorm2d with tracking_running_stats set to True'
        return ConvBn2d(conv, bn)
    return fuse_conv_bn_eval(conv, bn)


def fuse_conv_bn_relu(conv, bn, relu):
    assert (conv.training == bn.training == relu.training), "Conv and BN both must be in the same mode (train or eval)."
    if conv.training:
        assert bn.num_features == conv.out_channels, 'Output channel of Conv must match num_features of BatchNorm'
        assert bn.affine, 'Only support fusing BatchNorm with affine set to True'
        assert bn.track_running_stats, 'Only support fusing BatchNorm with tracking_running_stats set to True'
        if type(conv) is not Conv2d:
            raise NotImplementedError("Cannot fuse train modules: {}".format((conv, bn, relu)))
        return ConvBnReLU2d(conv, bn, relu)
    if type(conv) is not Conv2d:
        raise NotImplementedError("Cannot fuse eval modules: {}".format((conv, bn, relu)))
    return ConvReLU2d(fuse_conv_bn_eval(conv, bn), relu)
