"""Drop-in for src/models/stochastic/mcdropout/dropout.py."""
import torch
import torch.nn as nn

from ... import noise, ops


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

    def forward(self, x):
        # dropout.py:15-17: no self.training check; p <= 0 is the identity
        if self._p <= 0.0:
            return x
        # one keep/drop decision per (image, channel) for 4-D inputs, per element otherwise (dropout.py:19-30)
        mask = noise.pop_injected()
        key = (noise.seed(), self._qbn_layer_id, noise.next_draw())
        if x.dim() != 1:
            return ops.dropout_forward(x.detach(), self._p, mask, key)
        row_mask = None if mask is None else mask.unsqueeze(0)
        return ops.dropout_forward(x.detach().unsqueeze(0), self._p, row_mask, key)[0]

    def extra_repr(self):
        return 'p={}, quant={}'.format(self._p, False)
