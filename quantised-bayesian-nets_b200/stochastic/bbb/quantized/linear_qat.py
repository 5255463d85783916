"""Drop-in for src/models/stochastic/bbb/quantized/linear_qat.py: QAT Linear / LinearReLU with
fake-quantised mu, sigma, sampled weight and output (reference :18-41), from_float (:46-70)."""
import torch
import torch.nn.functional as F

from .... import config, noise, ops
from ..linear import Linear as LinearBBB
from ..linear import LinearReLU as LinearReLUBBB


def qat_eval_weight(mod, weight, std):
    """Eval branch shared by the QAT modules (linear_qat.py:30-37, conv_qat.py:40-48,154-162):
    W = FQ_add(w~ + FQ_mul(eps * sigma~)).  The FloatFunctionals carry the fake-quant observers."""
    eps = noise.pop_injected()
    if eps is None:
        seed, lid, draw = mod._key()
        eps = ops.philox_normal(weight.numel(), seed, lid, draw, device=weight.device).reshape(weight.shape)
    std_n = mod.mul_noise.mul(eps.to(weight.dtype), std)
    return mod.add_weight.add(weight, std_n)


class Linear(LinearBBB):
    _FLOAT_MODULE = LinearBBB

    def __init__(self, in_features, out_features, bias=False, qconfig=None, args=None):
        super(Linear, self).__init__(in_features, out_features, bias, args=args)
        assert qconfig, 'qconfig must be provided for QAT module'
        self.qconfig = qconfig
        self.weight_fake_quant = qconfig.weight()
        self.activation_post_process = qconfig.activation()
        self.std_fake_quant = qconfig.weight()

    def _forward(self, X):
        weight = self.weight_fake_quant(self.weight)
        std = self.std_fake_quant(F.softplus(self.std))
        if self.training:
            mode = config.pick_math_mode(self.in_features, self.out_features, lrt=True)
            eps = noise.pop_injected()
            return ops.LRTFunction.apply(X, weight, std, self.bias, 1, 0, 1, eps, self._key(), mode, True, None)
        w = qat_eval_weight(self, weight, std)
        d = ops.make_desc(X.shape[0], 1, 1, self.in_features, self.out_features, 1, 1)
        return ops.conv_forward(ops._f32(X).contiguous(), w.detach().contiguous().reshape(1, -1), d, 1, True, False, None, self.bias, None, False,
                                None, 1.0, ops.QBN_MATH_FP32)

    def forward(self, X):
        return self.activation_post_process(self._forward(X))

    def _get_name(self):
        return 'QATLinear'

    @classmethod
    def from_float(cls, mod, qconfig=None):
        assert type(mod) == cls._FLOAT_MODULE, ' qat.' + cls.__name__ + '.from_float only works for ' + cls._FLOAT_MODULE.__name__
        if not qconfig:
            assert hasattr(mod, 'qconfig'), 'Input float module must have qconfig defined'
            assert mod.qconfig, 'Input float module must have a valid qconfig'
        if type(mod) == LinearReLUBBB:
            mod = mod[0]
        qconfig = mod.qconfig
        q = cls(mod.in_features, mod.out_features, mod.bias is not None, qconfig)
        q.activation_post_process = mod.activation_post_process
        q.std_prior, q.weight, q.std, q.bias, q.args = mod.std_prior, mod.weight, mod.std, mod.bias, mod.args
        q.add_weight, q.mul_noise = mod.add_weight, mod.mul_noise
        q.add_weight.activation_post_process = qconfig.weight()
        q.mul_noise.activation_post_process = qconfig.weight()
        q._qbn_layer_id = mod._qbn_layer_id
        return q


class LinearReLU(Linear):
    _FLOAT_MODULE = LinearReLUBBB

    def forward(self, input):
        return self.activation_post_process(F.relu(self._forward(input)))

    @classmethod
    def from_float(cls, mod, qconfig=None):
        return super(LinearReLU, cls).from_float(mod, qconfig)

    def _get_name(self):
        return 'QATLinearReLU'
