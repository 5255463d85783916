"""Drop-in for src/models/stochastic/bbb/linear.py: same class names, constructor and forward
signatures, parameter names (`weight`=mu, `std`=rho, `std_prior`, `bias`) and FloatFunctional
attributes (`add_weight`, `mul_noise`), so reference checkpoints load both ways."""
import torch
import torch.nn as nn
from torch.nn import ReLU

from ... import config, noise, ops
from ._shared import attach_bayes_state, noise_key, typed_container
from .utils_bbb import kl_divergence_from_rho


class Linear(nn.Linear):
    def __init__(self, in_features, out_features, bias, sigma_prior=1.0, args=None):
        nn.Linear.__init__(self, in_features, out_features, bias)
        # linear.py:11-19: rho = -3 everywhere, prior std as a float32 [1] tensor, small uniform bias
        attach_bayes_state(self, -3.0, torch.ones((1,)) * sigma_prior, args)
        if self.bias is not None:
            with torch.no_grad():
                self.bias.uniform_(-0.01, 0.01)

    _key = noise_key

    def get_kl_divergence(self):
        """linear.py:24-28."""
        return kl_divergence_from_rho(self.weight, self.std, self.std_prior)

    def forward(self, x):
        vector = x.dim() == 1
        rows = x.unsqueeze(0) if vector else x
        if self.training:
            # linear.py:32-40 — LRT: both contractions + noise + bias in one kernel
            mode = config.pick_math_mode(self.in_features, self.out_features, lrt=True)
            out = ops.LRTFunction.apply(rows, self.weight, self.std, self.bias, 1, 0, 1, noise.pop_injected(), self._key(), mode, False, None)
        else:
            # linear.py:42-50 — one weight draw per forward, W = mu + softplus(rho)*eps
            out = eval_forward(self, rows, 1, 0, 1)
        return out[0] if vector else out


def _eval_forward_nograd(mod, x, stride, padding, dilation, relu=False, keep=None):
    xc = ops.nhwc(ops._f32(x))
    d = ops._geom(xc, mod.weight.shape, stride, padding, dilation)
    packed = ops.weight_prep(mod.weight, mod.std, False, None, want=("mu", "sigma"))
    eps = noise.pop_injected()
    eps_p = ops.pack_ohwi(ops._f32(eps)).reshape(1, -1) if eps is not None else None
    seed, lid, draw = mod._key()
    mode = config.pick_math_mode(d.C, d.N, lrt=False)
    w = ops.sample_weights(packed["mu"], packed["sigma"], 1, eps_p, seed, lid, draw, round_tf32=(mode == ops.QBN_MATH_TF32))
    if keep is not None:
        keep.update(xc=xc, d=d, mode=mode, packed=packed, eps_p=eps_p, key=(seed, lid, draw))
    return ops.conv_forward(xc, w, d, 1, True, False, None, mod.bias, None, relu, None, 1.0, mode)


class _EvalSampledFunction(torch.autograd.Function):
    """y = contract(x, mu + softplus(rho) * eps) + bias with the reference's gradients (linear.py:42-50 / conv.py:33-39 are
    plain autograd there): dx = g * W^T, dmu = g^T x, drho = dmu * eps * sigmoid(rho), dbias = sum g.  The contractions
    reuse the LRT backward kernels with the variance branch switched off (sigma^2 = 0, eps = 0)."""

    @staticmethod
    def forward(ctx, x, weight, std, bias, mod, stride, padding, dilation):
        keep = {}
        out = _eval_forward_nograd(mod, x, stride, padding, dilation, False, keep)
        n = weight.numel()
        eps_p = keep["eps_p"]
        if eps_p is None:                  # the sampler's own Philox stream: (seed, layer id, draw), counter = packed element / 4
            eps_p = ops.philox_normal(n, keep["key"][0], keep["key"][1], keep["key"][2], device=x.device)
        w_unrounded = keep["packed"]["mu"] + keep["packed"]["sigma"] * eps_p.reshape(-1)
        ctx.save_for_backward(keep["xc"], w_unrounded, eps_p.reshape(-1), std.detach())
        ctx.d, ctx.mode, ctx.wshape, ctx.has_bias = keep["d"], keep["mode"], tuple(weight.shape), bias is not None
        return out

    @staticmethod
    def backward(ctx, g):
        xc, w_p, eps_p, rho = ctx.saved_tensors
        gc = ops.nhwc(ops._f32(g))
        zero_w = torch.zeros_like(w_p)
        one = torch.ones_like(gc)
        dx, dw_p, _, dbias = ops.lrt_backward(xc, w_p, zero_w, gc, one, ctx.d, torch.zeros_like(gc), (0, 0, 0), ctx.needs_input_grad[0],
                                              ctx.has_bias, ctx.mode)
        def unpack(t):                     # packed OHWI -> the parameter's OIHW / [N, K] shape
            if len(ctx.wshape) == 2:
                return t.reshape(ctx.wshape)
            N, C, R, S = ctx.wshape
            return t.reshape(N, R, S, C).permute(0, 3, 1, 2).contiguous()
        d_mu = unpack(dw_p)
        d_rho = unpack(dw_p * eps_p) * torch.sigmoid(rho)
        return dx, d_mu, d_rho, dbias, None, None, None, None


def eval_forward(mod, x, stride, padding, dilation, relu=False):
    """Shared by Linear/Conv2d: sample W (A4) then contract; bias rides the epilogue.  Differentiable like the reference's
    eval-mode forward when autograd is recording (input or parameter requires grad); the no-grad product path is one sampler
    launch + one contraction launch."""
    if torch.is_grad_enabled() and (x.requires_grad or mod.weight.requires_grad or mod.std.requires_grad):
        out = _EvalSampledFunction.apply(x, mod.weight, mod.std, mod.bias, mod, stride, padding, dilation)
        return torch.relu(out) if relu else out
    with torch.no_grad():
        return _eval_forward_nograd(mod, x, stride, padding, dilation, relu)


LinearReLU = typed_container("LinearReLU", "Linear + ReLU awaiting QAT / conversion (linear.py:54-59).", Linear, ReLU, module=__name__)
