"""Drop-in for src/models/stochastic/mcdropout/dropout.py."""
import torch
import torch.nn as nn

from ... import noise, ops
from ...quant_utils import QTensor


def _frozen_scalar(value):
    return nn.Parameter(torch.ones((1,)) * value, requires_grad=False)


class BernoulliDropout(nn.Module):
    """Dropout that stays on at eval time (MC-Dropout).  The state-dict entries (`p`, `multiplier`) and FloatFunctional
    attributes are the reference's (dropout.py:9-13) so its checkpoints load; the mask itself is drawn and applied in
    one CUDA kernel (ops.dropout_forward), or taken from `noise.inject` when a test pins it."""

    def __init__(self, p=0.0):
        nn.Module.__init__(self)
        self._p = float(p)
        self.p = _frozen_scalar(p)
        self.multiplier = nn.Parameter(torch.ones((1,)) / (1.0 - self.p), requires_grad=False)
        for name in ("mul_mask", "mul_scalar"):
            setattr(self, name, torch.ao.nn.quantized.FloatFunctional())
        self._qbn_layer_id = noise.new_layer_id()

    def _prob(self):
        """`p` as the state-dict holds it now (a checkpoint may have replaced the constructor's value); cached per
        parameter version so the hot path never synchronises."""
        ver = (self.p._version, self.p.data_ptr())
        if self.__dict__.get("_p_ver") != ver:
            self._p, self._p_ver = float(self.p.detach().reshape(-1)[0]), ver
        return self._p

    def forward(self, x):
        # dropout.py:15-17: no self.training check; p <= 0 is the identity
        p = self._prob()
        if p <= 0.0:
            return x
        # one keep/drop decision per (image, channel) for 4-D inputs, per element otherwise (dropout.py:19-30)
        mask = noise.pop_injected()
        if isinstance(x, QTensor):
            return self._forward_int8(x, p, mask)
        key = (noise.seed(), self._qbn_layer_id, noise.next_draw())
        if x.dim() != 1:
            return ops.DropoutFunction.apply(x, p, mask, key)
        row_mask = None if mask is None else mask.unsqueeze(0)
        return ops.DropoutFunction.apply(x.unsqueeze(0), p, row_mask, key)[0]

    def _forward_int8(self, x, p, mask):
        """dropout.py:31-39 on quint8 activations: the mask is quantised at mul_mask's (scale, zero_point), quantized::mul
        writes the product at the same (scale, zero_point), mul_scalar leaves the integers and multiplies the scale."""
        fn = self.mul_mask
        if not (hasattr(fn, "scale") and hasattr(fn, "zero_point")):
            raise RuntimeError("BernoulliDropout got a quantised activation but mul_mask is not converted (quant_utils.convert)")
        s_m, z_m = float(fn.scale), int(fn.zero_point)
        sb = noise.sample_batch_state()
        q = x.q.contiguous(memory_format=torch.channels_last) if x.q.dim() == 4 else x.q.contiguous()
        if sb is not None and mask is None:
            # MC engine: activations carry n samples in the leading dimension; sample s draws Philox(seed, site, sample0 + s)
            n, s0, batch, bits = sb
            if q.shape[0] == batch:                      # input still shared by all samples: it forks here
                q = q.repeat(n, *([1] * (q.dim() - 1)))
                q = q.contiguous(memory_format=torch.channels_last) if q.dim() == 4 else q
            out = ops.i8_dropout_batched(q, x.scale, x.zero_point, p, s_m, z_m, n, (noise.seed(), self._qbn_layer_id, s0), act_bits=8)
        else:
            key = (noise.seed(), self._qbn_layer_id, noise.next_draw())
            out = ops.i8_dropout(q, x.scale, x.zero_point, p, s_m, z_m, mask, key, act_bits=8)
        mult = float(self.multiplier.detach().reshape(-1)[0]) if self.__dict__.get("_mult_ver") != self.multiplier._version else self._mult
        self._mult, self._mult_ver = mult, self.multiplier._version
        return QTensor(out, s_m * mult, z_m, 8)            # quantized::mul_scalar: scale * scalar in double, integers unchanged

    def extra_repr(self):
        return 'p={}, quant={}'.format(self._prob(), hasattr(self.mul_mask, 'zero_point'))
