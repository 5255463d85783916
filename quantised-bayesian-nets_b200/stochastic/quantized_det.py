"""Deterministic int8 conv / linear layers: the GPU twins of torch's stock nnq.Conv2d / nniq.ConvReLU2d / nnq.Linear /
nniq.LinearReLU, which is what the reference's MC-Dropout and SGHMC model families become after
`torch.quantization.convert` (src/quant_utils.py:140-141; models_mc.py:222, models_sgld.py:210) and what its SGHMC ensemble
evaluates (models_sgld.py:216-288).  Same kernels as the Bayesian int8 layers (qbn_i8_conv_fwd / qbn_i8_conv_p16_fwd, FBGEMM's
requantisation bit for bit) with one fixed weight tensor shared by every Monte-Carlo sample; per-tensor weight quantisation
(the reference's QConfig, quant_utils.py:129-138)."""
import torch
import torch.nn as nn

from .. import noise, ops
from ..quant_utils import QTensor


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


class _DetBase(nn.Module):
    RELU = False

    def bias(self):
        return self.bias_

    @staticmethod
    def _per_tensor(w):
        if w.qscheme() not in (torch.per_tensor_affine, torch.per_tensor_symmetric):
            raise NotImplementedError("per-channel weight quantisation is outside the reference's QConfig (quant_utils.py:133-138)")
        return w.int_repr().contiguous(), (float(w.q_scale()), int(w.q_zero_point()))

    @staticmethod
    def _batch(x):
        """(n_samples, per-sample batch, shared input?, activation width) of the surrounding noise.sample_batch, or a plain forward."""
        sb = noise.sample_batch_state()
        if sb is None:
            return 1, x.q.shape[0], True, 8
        n, _, batch, bits = sb
        rows = x.q.shape[0]
        if rows not in (batch, n * batch):
            raise ValueError("sample-batched forward: leading dimension %d is neither batch (%d) nor n_samples*batch (%d)" % (rows, batch, n * batch))
        return n, batch, rows == batch, bits


class QuantizedConv2d(_DetBase):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True, device="cuda"):
        super().__init__()
        if groups != 1:
            raise NotImplementedError("groups=1 only")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding, self.dilation, self.groups = _pair(kernel_size), _pair(stride), _pair(padding), _pair(dilation), 1
        self.weight = torch.zeros([out_channels, in_channels] + list(self.kernel_size), dtype=torch.int8, device=device)
        self.w_qp = (1.0, 0)
        self.bias_ = torch.zeros(out_channels, device=device) if bias else None
        self.scale, self.zero_point = 1.0, 0

    def _get_name(self):
        return "QuantizedConvReLU2d" if self.RELU else "QuantizedConv2d"

    def _apply(self, fn, *a, **k):
        self.weight = fn(self.weight)
        if self.bias_ is not None:
            self.bias_ = fn(self.bias_)
        return super()._apply(fn, *a, **k)

    def forward(self, x):
        assert isinstance(x, QTensor), "int8 modules take qbn_b200.quant_utils.QTensor activations"
        xq = x.q.contiguous(memory_format=torch.channels_last)
        _, C, H, W = xq.shape
        N, _, R, S = self.weight.shape
        n, batch, shared, bits = self._batch(x)
        wp = self.weight.permute(0, 2, 3, 1).contiguous().reshape(1, -1)        # packed OHWI, ONE tensor for all samples
        d = ops.make_desc(batch, H, W, C, N, R, S, self.stride, self.padding, self.dilation)
        y = ops.i8_conv_forward(xq, x.scale, x.zero_point, wp, self.w_qp[0], self.w_qp[1], d, self.bias_, self.scale, self.zero_point,
                                self.RELU, act_bits=bits, n_samples=n, x_shared=shared, w_shared=True, x_bits=x.bits)
        return QTensor(y, self.scale, self.zero_point, bits)

    @classmethod
    def from_torch(cls, mod):
        """From torch's nnq.Conv2d / nniq.ConvReLU2d (a CPU model converted by torch.quantization.convert)."""
        w, b = mod.weight(), mod.bias()
        q = cls(mod.in_channels, mod.out_channels, mod.kernel_size, mod.stride, mod.padding, mod.dilation, mod.groups, b is not None, device="cpu")
        q.weight, q.w_qp = cls._per_tensor(w)
        q.bias_ = None if b is None else b.detach().float().contiguous()
        q.scale, q.zero_point = float(mod.scale), int(mod.zero_point)
        return q


class QuantizedConvReLU2d(QuantizedConv2d):
    RELU = True


class QuantizedLinear(_DetBase):
    def __init__(self, in_features, out_features, bias=True, device="cuda"):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = torch.zeros((out_features, in_features), dtype=torch.int8, device=device)
        self.w_qp = (1.0, 0)
        self.bias_ = torch.zeros(out_features, device=device) if bias else None
        self.scale, self.zero_point = 1.0, 0

    def _get_name(self):
        return "QuantizedLinearReLU" if self.RELU else "QuantizedLinear"

    def _apply(self, fn, *a, **k):
        self.weight = fn(self.weight)
        if self.bias_ is not None:
            self.bias_ = fn(self.bias_)
        return super()._apply(fn, *a, **k)

    def forward(self, x):
        assert isinstance(x, QTensor), "int8 modules take qbn_b200.quant_utils.QTensor activations"
        xq = x.q.reshape(x.q.shape[0], -1).contiguous()
        n, batch, shared, bits = self._batch(x)
        d = ops.make_desc(batch, 1, 1, self.in_features, self.out_features, 1, 1)
        y = ops.i8_conv_forward(xq, x.scale, x.zero_point, self.weight.reshape(1, -1), self.w_qp[0], self.w_qp[1], d, self.bias_, self.scale,
                                self.zero_point, self.RELU, act_bits=bits, n_samples=n, x_shared=shared, w_shared=True, linear=True, x_bits=x.bits)
        return QTensor(y.reshape(-1, self.out_features), self.scale, self.zero_point, bits)

    @classmethod
    def from_torch(cls, mod):
        w, b = mod.weight(), mod.bias()
        q = cls(mod.in_features, mod.out_features, b is not None, device="cpu")
        q.weight, q.w_qp = cls._per_tensor(w)
        q.bias_ = None if b is None else b.detach().float().contiguous()
        q.scale, q.zero_point = float(mod.scale), int(mod.zero_point)
        return q


class QuantizedLinearReLU(QuantizedLinear):
    RELU = True
