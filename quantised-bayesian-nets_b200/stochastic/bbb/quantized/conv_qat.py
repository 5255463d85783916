"""Drop-in for src/models/stochastic/bbb/quantized/conv_qat.py: QAT Conv2d (:12-80), ConvBn2d with the
BN scale folded into mu and sigma before fake-quant and un-folded after the contraction (:139-167),
ConvBnReLU2d (:211-236), ConvReLU2d (:239-258), BN freeze/update API (:122-137)."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .... import config, noise, ops
from ..conv import Conv2d as Conv2dBBB
from ..conv import ConvBn2d as ConvBn2dBBB
from ..conv import ConvBnReLU2d as ConvBnReLU2dBBB
from ..conv import ConvReLU2d as ConvReLU2dBBB
from .linear_qat import qat_eval_weight


def _contract(mod, X, weight, std):
    """Train: LRT with the fake-quantised (mu~, sigma~).  Eval: sampled fake-quantised weight."""
    if mod.training:
        mode = config.pick_math_mode(mod.in_channels, mod.out_channels, lrt=True)
        eps = noise.pop_injected()
        return ops.LRTFunction.apply(X, weight, std, None, mod.stride, mod.padding, mod.dilation, eps, mod._key(), mode, True, None)
    w = qat_eval_weight(mod, weight, std)
    xc = ops.nhwc(ops._f32(X))
    d = ops._geom(xc, w.shape, mod.stride, mod.padding, mod.dilation)
    return ops.conv_forward(xc, ops.pack_ohwi(w.detach()).reshape(1, -1), d, 1, True, False, None, None, None, False, None, 1.0, ops.QBN_MATH_FP32)


class Conv2d(Conv2dBBB):
    _FLOAT_MODULE = Conv2dBBB

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=False,
                 padding_mode='zeros', qconfig=None, args=None):
        super(Conv2d, self).__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, padding_mode, args=args)
        assert qconfig, 'qconfig must be provided for QAT module'
        self.qconfig = qconfig
        self.weight_fake_quant = qconfig.weight()
        self.activation_post_process = qconfig.activation()
        self.std_fake_quant = qconfig.weight()

    def _forward(self, X):
        weight = self.weight_fake_quant(self.weight)
        std = self.std_fake_quant(F.softplus(self.std))
        Z = _contract(self, X, weight, std)
        if self.bias is not None:
            Z = Z + self.bias.reshape(1, -1, 1, 1)
        return Z

    def forward(self, input):
        return self.activation_post_process(self._forward(input))

    def _get_name(self):
        return 'QATConv2d'

    @classmethod
    def from_float(cls, mod, qconfig=None):
        assert type(mod) == cls._FLOAT_MODULE, ' qat.' + cls.__name__ + '.from_float only works for ' + cls._FLOAT_MODULE.__name__
        if not qconfig:
            assert hasattr(mod, 'qconfig'), 'Input float module must have qconfig defined'
            assert mod.qconfig, 'Input float module must have a valid qconfig'
        if isinstance(mod, ConvReLU2dBBB):
            mod = mod[0]
        qconfig = mod.qconfig
        q = cls(mod.in_channels, mod.out_channels, mod.kernel_size, mod.stride, mod.padding, mod.dilation, mod.groups,
                mod.bias is not None, mod.padding_mode, qconfig)
        q.activation_post_process = mod.activation_post_process
        q.weight, q.std, q.std_prior, q.bias, q.args = mod.weight, mod.std, mod.std_prior, mod.bias, mod.args
        q.add_weight, q.mul_noise = mod.add_weight, mod.mul_noise
        q.add_weight.activation_post_process = qconfig.weight()
        q.mul_noise.activation_post_process = qconfig.weight()
        q._qbn_layer_id = mod._qbn_layer_id
        return q


class ConvBn2d(Conv2d):
    _version = 1
    _FLOAT_MODULE = ConvBn2dBBB

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=False,
                 padding_mode='zeros', eps=1e-05, momentum=0.1, freeze_bn=False, qconfig=None, args=None):
        super(ConvBn2d, self).__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias,
                                       padding_mode, qconfig, args)
        self.freeze_bn = freeze_bn if self.training else True
        self.bn = nn.BatchNorm2d(out_channels, eps, momentum, True, True)
        self.reset_bn_parameters()
        if self.training:
            if freeze_bn:
                self.freeze_bn_stats()
            else:
                self.update_bn_stats()
        else:
            self.freeze_bn_stats()

    def reset_running_stats(self):
        self.bn.reset_running_stats()

    def reset_bn_parameters(self):
        self.bn.reset_running_stats()
        torch.nn.init.uniform_(self.bn.weight)
        torch.nn.init.zeros_(self.bn.bias)
        if self.bias is not None:
            fan_in, _ = torch.nn.init._calculate_fan_in_and_fan_out(self.weight)
            bound = 1 / math.sqrt(fan_in)
            torch.nn.init.uniform_(self.bias, -bound, bound)

    def update_bn_stats(self):
        self.freeze_bn = False
        self.bn.training = True
        return self

    def freeze_bn_stats(self):
        self.freeze_bn = True
        self.bn.training = False
        return self

    def train(self, mode=True):
        self.training = mode
        if not self.freeze_bn:
            for module in self.children():
                module.train(mode)
        return self

    def _forward(self, X):
        # conv_qat.py:139-167: scale mu and sigma by gamma/sqrt(running_var+eps) BEFORE fake-quant,
        # contract, divide the scale back out, add the conv bias, run the real BatchNorm.
        running_std = torch.sqrt(self.bn.running_var + self.bn.eps)
        scale_factor = self.bn.weight / running_std
        weight = self.weight_fake_quant(self.weight * scale_factor.reshape([-1, 1, 1, 1]))
        std = self.std_fake_quant(F.softplus(self.std) * scale_factor.reshape([-1, 1, 1, 1]))
        Z = _contract(self, X, weight, std)
        Z_orig = Z / scale_factor.reshape([1, -1, 1, 1])
        if self.bias is not None:
            Z_orig = Z_orig + self.bias.reshape([1, -1, 1, 1])
        return self.bn(Z_orig)

    def _get_name(self):
        return 'QATConvBn2d'

    @classmethod
    def from_float(cls, mod, qconfig=None):
        assert type(mod) == cls._FLOAT_MODULE, 'qat.' + cls.__name__ + '.from_float only works for ' + cls._FLOAT_MODULE.__name__
        conv, bn = mod[0], mod[1]
        if not qconfig:
            qconfig = getattr(mod, 'qconfig', None) or conv.qconfig
        q = cls(conv.in_channels, conv.out_channels, conv.kernel_size, conv.stride, conv.padding, conv.dilation, conv.groups,
                conv.bias is not None, conv.padding_mode, bn.eps, bn.momentum, False, qconfig)
        q.activation_post_process = conv.activation_post_process
        q.mul_noise, q.add_weight = conv.mul_noise, conv.add_weight
        q.mul_noise.activation_post_process = qconfig.weight()
        q.add_weight.activation_post_process = qconfig.weight()
        q.weight, q.std, q.std_prior, q.bias, q.args = conv.weight, conv.std, conv.std_prior, conv.bias, conv.args
        q.bn.weight, q.bn.bias = bn.weight, bn.bias
        q.bn.running_mean, q.bn.running_var, q.bn.num_batches_tracked = bn.running_mean, bn.running_var, bn.num_batches_tracked
        q._qbn_layer_id = conv._qbn_layer_id
        return q


class ConvBnReLU2d(ConvBn2d):
    _FLOAT_MODULE = ConvBnReLU2dBBB

    def forward(self, input):
        return self.activation_post_process(F.relu(self._forward(input)))

    def _get_name(self):
        return 'QATConvBnReLU2d'


class ConvReLU2d(Conv2d):
    _FLOAT_MODULE = ConvReLU2dBBB

    def forward(self, input):
        return self.activation_post_process(F.relu(self._forward(input)))

    def _get_name(self):
        return 'QATConvReLU2d'
