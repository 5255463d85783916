"""Drop-in for src/models/stochastic/bbb/quantized/linear_q.py: true-int8 BBB Linear / LinearReLU.
Per forward (reference :80-94,154-173): eps -> int8 @ NOISE_SCALE -> qmul(sigma_q, eps_q) ->
qadd(mu_q, .) -> clamp_weight -> u8 x s8 -> s32 contraction -> FBGEMM requantisation — here as two
CUDA launches (qbn_i8_sample_weights, qbn_i8_conv_fwd), bit-exact against the reference's CPU path."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .... import noise, ops
from ....quant_utils import INT_BOUNDS, QFunctional, QTensor
from . import NOISE_SCALE, NOISE_ZERO_POINT


def functional_qparams(ff):
    """(scale, zero_point) a QFunctional.from_float would take from the FloatFunctional's observer."""
    if hasattr(ff, "scale") and hasattr(ff, "zero_point") and not hasattr(ff.activation_post_process, "calculate_qparams"):
        return float(ff.scale), int(ff.zero_point)
    s, z = ff.activation_post_process.calculate_qparams()
    return float(s), int(z)


def quantise_param(t, fq):
    """`weight_post_process(w)` then `_quantize_weight(w, weight_post_process)` (linear_q.py:124-132):
    one more observer update on the tensor, then quantize_per_tensor with its qparams."""
    fq(t.detach())
    s, z = fq.calculate_qparams()
    return ops.quantize_s8(t.detach(), float(s), int(z), -128, 127), float(s), int(z)


class _I8Base(nn.Module):
    RELU = False

    def _init_common(self, args):
        self.args = args
        self.scale, self.zero_point = 1.0, 0
        self.bias_ = None
        self.std_prior = torch.nn.Parameter(torch.ones((1,)), requires_grad=False)
        # the two QFunctionals of the reference module (linear_q.py:30-31, conv_q.py:62-63): they hold the output
        # qparams of sigma_q*eps_q and of mu_q + (.), and give the checkpoint its `add_weight.*` / `mul_noise.*` keys
        self.add_weight, self.mul_noise = QFunctional(), QFunctional()
        self._qbn_layer_id = noise.new_layer_id()

    @property
    def mul_qp(self):
        return (self.mul_noise.scale, self.mul_noise.zero_point)

    @mul_qp.setter
    def mul_qp(self, qp):
        self.mul_noise.scale, self.mul_noise.zero_point = float(qp[0]), int(qp[1])

    @property
    def add_qp(self):
        return (self.add_weight.scale, self.add_weight.zero_point)

    @add_qp.setter
    def add_qp(self, qp):
        self.add_weight.scale, self.add_weight.zero_point = float(qp[0]), int(qp[1])

    def bias(self):
        return self.bias_

    def _sample_params(self):
        p = ops.I8SampleParams()
        p.s_mu, p.z_mu, p.s_sigma, p.z_sigma = self.mu_qp[0], self.mu_qp[1], self.sigma_qp[0], self.sigma_qp[1]
        p.s_eps, p.z_eps = NOISE_SCALE, NOISE_ZERO_POINT
        p.s_mul, p.z_mul, p.s_add, p.z_add = self.mul_qp[0], self.mul_qp[1], self.add_qp[0], self.add_qp[1]
        wbits = getattr(self.args, "weight_precision", 8) if self.args is not None else 8
        p.w_min, p.w_max = INT_BOUNDS[wbits]
        p.n_vec = -1
        return p

    def sampled_weight(self):
        """int8 OIHW sampled weight of this forward (steps 1-4 of SURVEY §8a row A6)."""
        eps = noise.pop_injected()
        key = (noise.seed(), self._qbn_layer_id, noise.next_draw())
        n = self.weight.numel()
        w = ops.i8_sample_weights(self.weight.reshape(-1), self.std.reshape(-1), self._sample_params(), 1,
                                  eps.float().reshape(1, -1).contiguous() if eps is not None else None, key[0], key[1], key[2])
        return w.reshape(self.weight.shape)

    def sampled_weights(self, n_samples, sample0):
        """[n_samples, *weight.shape] int8: the draws of global samples sample0.. (same stream as `noise.sample_index(s)`)."""
        w = ops.i8_sample_weights(self.weight.reshape(-1), self.std.reshape(-1), self._sample_params(), n_samples, None,
                                  noise.seed(), self._qbn_layer_id, sample0)
        return w.reshape((n_samples,) + tuple(self.weight.shape))

    @staticmethod
    def _batch_layout(x, n_samples, batch):
        """Is the activation shared by the samples ([batch, ...]) or already per sample ([n_samples*batch, ...])?"""
        rows = x.q.shape[0]
        if rows != batch and rows != n_samples * batch:
            raise ValueError("sample-batched forward: leading dimension %d is neither batch (%d) nor n_samples*batch (%d)"
                             % (rows, batch, n_samples * batch))
        return rows == batch

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        super()._save_to_state_dict(destination, prefix, keep_vars)
        destination[prefix + 'scale'] = torch.tensor(self.scale)
        destination[prefix + 'zero_point'] = torch.tensor(self.zero_point)
        # same keys as the reference (linear_q.py:40-46): weight/std as torch per-tensor-affine qint8 tensors
        destination[prefix + 'weight'] = torch._make_per_tensor_quantized_tensor(self.weight.cpu(), self.mu_qp[0], self.mu_qp[1])
        destination[prefix + 'std'] = torch._make_per_tensor_quantized_tensor(self.std.cpu(), self.sigma_qp[0], self.sigma_qp[1])
        destination[prefix + 'bias_'] = self.bias_

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        dev = self.weight.device
        self.scale = float(state_dict.pop(prefix + 'scale'))
        self.zero_point = int(state_dict.pop(prefix + 'zero_point'))
        w = state_dict.pop(prefix + 'weight')
        s = state_dict.pop(prefix + 'std')
        self.weight, self.mu_qp = w.int_repr().to(dev), (float(w.q_scale()), int(w.q_zero_point()))
        self.std, self.sigma_qp = s.int_repr().to(dev), (float(s.q_scale()), int(s.q_zero_point()))
        b = state_dict.pop(prefix + 'bias_')
        self.bias_ = None if b is None else b.detach().to(device=dev, dtype=torch.float32).contiguous()
        super()._load_from_state_dict(state_dict, prefix, local_metadata, False, missing_keys, unexpected_keys, error_msgs)


class Linear(_I8Base):
    _version = 1

    def __init__(self, in_features, out_features, bias_=False, args=None, device="cuda"):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self._init_common(args)
        self.weight = torch.zeros((out_features, in_features), dtype=torch.int8, device=device)
        self.std = torch.zeros((out_features, in_features), dtype=torch.int8, device=device)
        self.mu_qp, self.sigma_qp, self.mul_qp, self.add_qp = (1.0, 0), (1.0, 0), (1.0, 0), (1.0, 0)
        if bias_:
            self.bias_ = torch.zeros(out_features, device=device)

    def _get_name(self):
        return 'QuantizedLinear'

    def extra_repr(self):
        return 'in_features={}, out_features={}, scale={}, zero_point={}, bias={}'.format(
            self.in_features, self.out_features, self.scale, self.zero_point, self.bias() is not None)

    def forward(self, x):
        assert isinstance(x, QTensor), "int8 modules take qbn_b200.quant_utils.QTensor activations"
        xq = x.q.reshape(x.q.shape[0], -1).contiguous()
        sb = noise.sample_batch_state()
        if sb is not None:                                          # all samples of an MC chunk in one launch
            n, s0, batch, bits = sb
            shared = self._batch_layout(x, n, batch)
            w = self.sampled_weights(n, s0).reshape(n, -1)
            d = ops.make_desc(batch, 1, 1, self.in_features, self.out_features, 1, 1)
            y = ops.i8_conv_forward(xq, x.scale, x.zero_point, w, self.add_qp[0], self.add_qp[1], d, self.bias(), self.scale,
                                    self.zero_point, self.RELU, act_bits=bits, n_samples=n, x_shared=shared, linear=True, x_bits=x.bits)
            return QTensor(y.reshape(-1, self.out_features), self.scale, self.zero_point, bits)
        w = self.sampled_weight()
        d = ops.make_desc(xq.shape[0], 1, 1, self.in_features, self.out_features, 1, 1)
        y = ops.i8_conv_forward(xq, x.scale, x.zero_point, w.reshape(1, -1), self.add_qp[0], self.add_qp[1], d, self.bias(),
                                self.scale, self.zero_point, self.RELU, act_bits=8, linear=True, x_bits=x.bits)
        return QTensor(y, self.scale, self.zero_point)

    @classmethod
    def from_float(cls, mod):
        assert hasattr(mod, 'weight_fake_quant'), "convert from the QAT module (prepare_model first)"
        dev = mod.weight.device
        q = cls(mod.in_features, mod.out_features, mod.bias is not None, args=mod.args, device=dev)
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


class LinearReLU(Linear):
    RELU = True

    def _get_name(self):
        return 'QuantizedLinearReLU'
