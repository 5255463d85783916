"""Mirror of src/models/stochastic/bbb/utils_bbb.py (reference :3-8)."""
import torch

from ... import ops


def kl_divergence(mu, sigma_or_rho, mu_prior=None, sigma_prior=1.0, from_rho=True):
    """Closed-form Gaussian KL (utils_bbb.py:3-5) with mu_prior = 0 and a scalar sigma_prior, as
    every reference call site uses it (linear.py:24-28, conv.py:43-47).  One fused CUDA pass that
    also produces the gradient (ops.KLFunction); takes rho (sigma = softplus(rho))."""
    if not from_rho:
        raise NotImplementedError("kl_divergence takes rho; the reference always passes softplus(self.std)")
    sp = float(sigma_prior.reshape(-1)[0]) if torch.is_tensor(sigma_prior) else float(sigma_prior)
    return ops.kl_divergence(mu, sigma_or_rho, sp)


def softplusinv(x):
    """utils_bbb.py:7-8 (weight-sized, used once at BN-fold time: conv.py:77)."""
    return torch.log(torch.exp(x) - 1.)
