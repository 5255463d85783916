"""Drop-in for src/models/stochastic/mcdropout/dropout.py."""
import torch
import torch.nn as nn

from ... import noise, ops


class BernoulliDropout(nn.Module):
    def __init__(self, p=0.0):
        super(BernoulliDropout, self).__init__()
        self.p = torch.nn.Parameter(torch.ones((1,)) * p, requires_grad=False)                       # dropout.py:9
        self.multiplier = torch.nn.Parameter(torch.ones((1,)) / (1.0 - self.p), requires_grad=False)  # dropout.py:10
        self.mul_mask = torch.ao.nn.quantized.FloatFunctional()
        self.mul_scalar = torch.ao.nn.quantized.FloatFunctional()
        self._qbn_layer_id = noise.new_layer_id()
        self._p = float(p)

    def forward(self, x):
        # dropout.py:15-17: ALWAYS active (no self.training check), identity only when p <= 0
        if self._p <= 0.0:
            return x
        mask = noise.pop_injected()  # [B,C] (4-D input) or x.shape (<=2-D), dropout.py:19-30
        key = (noise.seed(), self._qbn_layer_id, noise.next_draw())
        squeeze = x.dim() == 1
        if squeeze:
            x, mask = x.unsqueeze(0), (mask.unsqueeze(0) if mask is not None else None)
        out = ops.dropout_forward(x.detach(), self._p, mask, key)
        return out.squeeze(0) if squeeze else out

    def extra_repr(self):
        return 'p={}, quant={}'.format(self._p, False)
