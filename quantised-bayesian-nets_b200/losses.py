"""Host-side mirror of src/losses.py (the ELBO the trainer builds around the hot path, trainer.py:96-104): same factory,
class names and call signature `loss(output, target, kl, gamma, n_batches, n_points) -> (loss, data_term, kl_term)`.

Classification on CUDA tensors: ONE launch (`qbn_elbo_cls`) produces the loss, its two terms and d loss / d probs — the KL enters
as the 0-d device tensor of `qbn_kl_multi` (one launch for all layers), nothing is read back to the host.  Regression, and any
input that does not live on a CUDA device (CPU tests of the mirror), take the reference's torch expressions."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class Loss(nn.Module):
    def __init__(self, args, scaling):
        super().__init__()
        self.args, self.scaling = args, scaling

    def _scaled(self, data_term, kl, batch, n_batches, n_points):
        """losses.py:19-25: 'whole' = data term of the whole training set vs KL per batch; 'batch' = mean data term vs KL per
        example."""
        if self.scaling == "whole":
            return n_points * data_term * self.args.loss_multiplier, kl / n_batches
        if self.scaling == "batch":
            return data_term, kl / (batch * n_batches)
        raise NotImplementedError("Other scaling not implemented!")               # losses.py:26-27,49-50

    def forward(self, output, target, kl, gamma, n_batches, n_points):
        data_term, kl_term = self._scaled(self.data_term(output, target), kl, target.shape[0], n_batches, n_points)
        return data_term + gamma * kl_term, data_term, kl_term


class _ElboCls(torch.autograd.Function):
    """losses.py:14-29 fused: returns (loss, data term, kl term) as 0-d tensors; backward = two scalar-times-tensor products."""

    @staticmethod
    def forward(ctx, output, target, kl, data_scale, kl_scale, gamma):
        import ctypes
        from . import _lib
        out_c = output.detach().float().contiguous()
        B, K = out_c.shape
        res = torch.empty(3, dtype=torch.float32, device=out_c.device)
        d_probs = torch.empty_like(out_c) if output.requires_grad else None
        klc = kl.detach().float().reshape(1).contiguous()
        p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
        _lib.call("qbn_elbo_cls", p(out_c), p(target.contiguous()), p(klc), B, K, float(data_scale), float(kl_scale), float(gamma), p(res), p(d_probs),
                  ctypes.c_void_p(torch.cuda.current_stream(out_c.device).cuda_stream))
        ctx.save_for_backward(d_probs)
        ctx.kl_scale, ctx.gamma, ctx.kl_shape = float(kl_scale), float(gamma), kl.shape
        return res[0], res[1], res[2]

    @staticmethod
    def backward(ctx, g_loss, g_data, g_kl):
        (d_probs,) = ctx.saved_tensors
        d_out = (g_loss + g_data) * d_probs if d_probs is not None else None
        d_kl = ((g_loss * ctx.gamma + g_kl) * ctx.kl_scale).reshape(ctx.kl_shape)
        return d_out, None, d_kl, None, None, None


class ClassificationLoss(Loss):
    """NLL of the (already soft-maxed) model output, with the reference's 1e-8 guard (losses.py:14-29)."""

    def data_term(self, output, target):
        return F.nll_loss(torch.log(output + 1e-8), target)

    def forward(self, output, target, kl, gamma, n_batches, n_points):
        fused = (torch.is_tensor(output) and output.is_cuda and output.dim() == 2 and torch.is_tensor(kl) and kl.is_cuda and kl.numel() == 1
                 and target.dtype == torch.int64 and target.is_cuda)
        if not fused:
            return super().forward(output, target, kl, gamma, n_batches, n_points)
        if self.scaling == "whole":
            data_scale, kl_scale = n_points * self.args.loss_multiplier, 1.0 / n_batches
        elif self.scaling == "batch":
            data_scale, kl_scale = 1.0, 1.0 / (target.shape[0] * n_batches)
        else:
            raise NotImplementedError("Other scaling not implemented!")           # losses.py:26-27
        return _ElboCls.apply(output, target, kl, data_scale, kl_scale, gamma)


class RegressionLoss(Loss):
    """Heteroscedastic Gaussian NLL on output = (mean, var): mean_b sum_d ((t-m)^2/(var+1e-8) + log(var+1e-8))
    (losses.py:31-52)."""

    def data_term(self, output, target):
        mean, var = output[0], output[1]
        return torch.mean(torch.sum((target - mean) ** 2 / (var + 1e-8) + torch.log(var + 1e-8), 1), 0)


LOSS_FACTORY = {"classification": lambda args, scaling: ClassificationLoss(args, scaling),
                "regression": lambda args, scaling: RegressionLoss(args, scaling)}
