"""Drop-in for src/models/stochastic/bbb/quantized/linear_qat.py: QAT Linear / LinearReLU with
fake-quantised mu, sigma, sampled weight and output (reference :18-41), from_float (:46-70)."""
from .... import config, noise, ops
from .._shared import QATMixin
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


class Linear(QATMixin, LinearBBB):
    _FLOAT_MODULE = LinearBBB
    _NAME = 'QATLinear'

    def __init__(self, in_features, out_features, bias=False, qconfig=None, args=None):
        LinearBBB.__init__(self, in_features, out_features, bias, args=args)
        self._attach_qat(qconfig)

    def _forward(self, X):
        weight, std = self._fake_quantised()
        if self.training:
            mode = config.pick_math_mode(self.in_features, self.out_features, lrt=True)
            return ops.LRTFunction.apply(X, weight, std, self.bias, 1, 0, 1, noise.pop_injected(), self._key(), mode, True, None)
        w = qat_eval_weight(self, weight, std).detach().contiguous().reshape(1, -1)
        d = ops.make_desc(X.shape[0], 1, 1, self.in_features, self.out_features, 1, 1)
        return ops.conv_forward(ops._f32(X).contiguous(), w, d, 1, True, False, None, self.bias, None, False, None, 1.0, ops.QBN_MATH_FP32)

    @classmethod
    def from_float(cls, mod, qconfig=None):
        cls._check_float(mod, qconfig)
        src = mod[0] if type(mod) == LinearReLUBBB else mod
        return cls._adopt(cls(src.in_features, src.out_features, src.bias is not None, src.qconfig), src, src.qconfig)


class LinearReLU(Linear):
    _FLOAT_MODULE = LinearReLUBBB
    _NAME = 'QATLinearReLU'
    _RELU = True
