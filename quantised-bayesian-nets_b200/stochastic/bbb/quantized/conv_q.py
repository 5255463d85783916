"""Drop-in for src/models/stochastic/bbb/quantized/conv_q.py: true-int8 BBB Conv2d / ConvReLU2d
(reference :107-125,189-209), from_float incl. the BatchNorm fold of QAT ConvBn modules (:127-177)."""
import torch
import torch.nn.functional as F

from .... import noise, ops
from ....quant_utils import QTensor
from ..conv import fuse_conv_bn_weights
from .conv_qat import ConvBn2d as ConvBn2dQAT
from .linear_q import _I8Base, functional_qparams, quantise_param


class Conv2d(_I8Base):
    _version = 1

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=False,
                 padding_mode='zeros', args=None, device="cuda"):
        super().__init__()
        if padding_mode != 'zeros':
            raise NotImplementedError("Currently only zero-padding is supported by quantized conv")
        if groups != 1:
            raise NotImplementedError("groups=1 only")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding, self.dilation, self.groups = kernel_size, stride, padding, dilation, groups
        self.padding_mode = padding_mode
        self._init_common(args)
        shape = [out_channels, in_channels] + list(kernel_size)
        self.weight = torch.zeros(shape, dtype=torch.int8, device=device)
        self.std = torch.zeros(shape, dtype=torch.int8, device=device)
        self.mu_qp, self.sigma_qp, self.mul_qp, self.add_qp = (1.0, 0), (1.0, 0), (1.0, 0), (1.0, 0)
        if bias:
            self.bias_ = torch.zeros(out_channels, device=device)

    def _get_name(self):
        return 'QuantizedConv2d'

    def forward(self, x):
        assert isinstance(x, QTensor), "int8 modules take qbn_b200.quant_utils.QTensor activations"
        if x.q.dim() != 4:
            raise ValueError("Input shape must be `(N, C, H, W)`!")
        xq = x.q.contiguous(memory_format=torch.channels_last)
        B, C, H, W = xq.shape
        N, _, R, S = self.weight.shape
        sb = noise.sample_batch_state()
        if sb is not None:                                          # all samples of an MC chunk in one launch
            n, s0, batch, bits = sb
            shared = self._batch_layout(x, n, batch)
            wp = self.sampled_weights(n, s0).permute(0, 1, 3, 4, 2).contiguous().reshape(n, -1)     # [n][OHWI]
            d = ops.make_desc(batch, H, W, C, N, R, S, self.stride, self.padding, self.dilation)
            y = ops.i8_conv_forward(xq, x.scale, x.zero_point, wp, self.add_qp[0], self.add_qp[1], d, self.bias(), self.scale,
                                    self.zero_point, self.RELU, act_bits=bits, n_samples=n, x_shared=shared, x_bits=x.bits)
            return QTensor(y, self.scale, self.zero_point, bits)
        w = self.sampled_weight()                                   # OIHW int8
        wp = w.permute(0, 2, 3, 1).contiguous().reshape(1, -1)      # packed OHWI
        d = ops.make_desc(B, H, W, C, N, R, S, self.stride, self.padding, self.dilation)
        y = ops.i8_conv_forward(xq, x.scale, x.zero_point, wp, self.add_qp[0], self.add_qp[1], d, self.bias(), self.scale, self.zero_point,
                                self.RELU, act_bits=8, x_bits=x.bits)
        return QTensor(y, self.scale, self.zero_point)

    @classmethod
    def from_float(cls, mod):
        assert hasattr(mod, 'weight_fake_quant'), "convert from the QAT module (prepare_model first)"
        if isinstance(mod, ConvBn2dQAT):                            # conv_q.py:130-133,216-219
            mod.weight, mod.bias, mod.std = fuse_conv_bn_weights(mod.weight, mod.bias, mod.std, mod.bn.running_mean, mod.bn.running_var,
                                                                mod.bn.eps, mod.bn.weight, mod.bn.bias)
        dev = mod.weight.device
        q = cls(mod.in_channels, mod.out_channels, mod.kernel_size, mod.stride, mod.padding, mod.dilation, mod.groups,
                mod.bias is not None, mod.padding_mode, args=mod.args, device=dev)
        q.weight, s, z = quantise_param(mod.weight.float(), mod.weight_fake_quant)
        q.mu_qp = (s, z)
        q.std, s, z = quantise_param(F.softplus(mod.std.float()), mod.std_fake_quant)
        q.sigma_qp = (s, z)
        a_s, a_z = mod.activation_post_process.calculate_qparams()
        q.scale, q.zero_point = float(a_s), int(a_z)
        q.mul_qp, q.add_qp = functional_qparams(mod.mul_noise), functional_qparams(mod.add_weight)
        q.std_prior = mod.std_prior
        q.bias_ = mod.bias.detach() if mod.bias is not None else None
        q._qbn_layer_id = mod._qbn_layer_id
        return q


class ConvReLU2d(Conv2d):
    RELU = True

    def _get_name(self):
        return 'QuantizedConvReLU2d'
