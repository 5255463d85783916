"""Drop-in for src/models/stochastic/bbb/quantized/conv_qat.py: QAT Conv2d (:12-80), ConvBn2d with the
BN scale folded into mu and sigma before fake-quant and un-folded after the contraction (:139-167),
ConvBnReLU2d (:211-236), ConvReLU2d (:239-258), BN freeze/update API (:122-137)."""
import math

import torch
import torch.nn as nn

from .... import config, noise, ops
from .._shared import QATMixin
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


def _per_channel(v):
    return v.reshape(1, -1, 1, 1)


def _conv_ctor_args(conv):
    return (conv.in_channels, conv.out_channels, conv.kernel_size, conv.stride, conv.padding, conv.dilation, conv.groups,
            conv.bias is not None, conv.padding_mode)


class Conv2d(QATMixin, Conv2dBBB):
    _FLOAT_MODULE = Conv2dBBB
    _NAME = 'QATConv2d'

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=False,
                 padding_mode='zeros', qconfig=None, args=None):
        Conv2dBBB.__init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, padding_mode, args=args)
        self._attach_qat(qconfig)

    def _forward(self, X):
        out = _contract(self, X, *self._fake_quantised())
        return out if self.bias is None else out + _per_channel(self.bias)

    @classmethod
    def from_float(cls, mod, qconfig=None):
        cls._check_float(mod, qconfig)
        src = mod[0] if isinstance(mod, ConvReLU2dBBB) else mod
        return cls._adopt(cls(*_conv_ctor_args(src), src.qconfig), src, src.qconfig)


class ConvBn2d(Conv2d):
    _version = 1
    _FLOAT_MODULE = ConvBn2dBBB
    _NAME = 'QATConvBn2d'

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=False,
                 padding_mode='zeros', eps=1e-05, momentum=0.1, freeze_bn=False, qconfig=None, args=None):
        Conv2d.__init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, padding_mode, qconfig, args)
        self.bn = nn.BatchNorm2d(out_channels, eps, momentum, True, True)
        self.reset_bn_parameters()
        # conv_qat.py:97-106: statistics move only in training mode and only when not frozen
        self.freeze_bn = bool(freeze_bn) or not self.training
        (self.freeze_bn_stats if self.freeze_bn else self.update_bn_stats)()

    def reset_running_stats(self):
        self.bn.reset_running_stats()

    def reset_bn_parameters(self):
        """conv_qat.py:111-120: gamma ~ U(0,1), beta = 0, conv bias ~ U(+-1/sqrt(fan_in))."""
        self.reset_running_stats()
        nn.init.uniform_(self.bn.weight)
        nn.init.zeros_(self.bn.bias)
        if self.bias is not None:
            bound = 1 / math.sqrt(nn.init._calculate_fan_in_and_fan_out(self.weight)[0])
            nn.init.uniform_(self.bias, -bound, bound)

    def _set_bn_frozen(self, frozen):
        self.freeze_bn, self.bn.training = frozen, not frozen
        return self

    def update_bn_stats(self):
        return self._set_bn_frozen(False)

    def freeze_bn_stats(self):
        return self._set_bn_frozen(True)

    def train(self, mode=True):
        """conv_qat.py:131-137: a frozen BatchNorm stays in eval mode whatever the parent does."""
        self.training = mode
        for child in (() if self.freeze_bn else self.children()):
            child.train(mode)
        return self

    def _forward(self, X):
        # conv_qat.py:139-167: scale mu and sigma by gamma/sqrt(running_var+eps) BEFORE fake-quant,
        # contract, divide the scale back out, add the conv bias, run the real BatchNorm.
        scale = self.bn.weight / torch.sqrt(self.bn.running_var + self.bn.eps)
        out = _contract(self, X, *self._fake_quantised(scale.reshape(-1, 1, 1, 1))) / _per_channel(scale)
        return self.bn(out if self.bias is None else out + _per_channel(self.bias))

    @classmethod
    def from_float(cls, mod, qconfig=None):
        if type(mod) != cls._FLOAT_MODULE:
            raise AssertionError('qat.' + cls.__name__ + '.from_float only works for ' + cls._FLOAT_MODULE.__name__)
        conv, bn = mod[0], mod[1]
        qconfig = qconfig or getattr(mod, 'qconfig', None) or conv.qconfig
        q = cls._adopt(cls(*_conv_ctor_args(conv), bn.eps, bn.momentum, False, qconfig), conv, qconfig)
        for name in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked"):
            setattr(q.bn, name, getattr(bn, name))
        return q


class ConvBnReLU2d(ConvBn2d):
    _FLOAT_MODULE = ConvBnReLU2dBBB
    _NAME = 'QATConvBnReLU2d'
    _RELU = True


class ConvReLU2d(Conv2d):
    _FLOAT_MODULE = ConvReLU2dBBB
    _NAME = 'QATConvReLU2d'
    _RELU = True
