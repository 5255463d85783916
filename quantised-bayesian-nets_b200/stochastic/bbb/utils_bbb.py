"""Mirror of src/models/stochastic/bbb/utils_bbb.py (reference :3-8)."""
import torch

from ... import ops


def kl_divergence(mu, sigma_or_rho, mu_prior=None, sigma_prior=1.0, from_rho=True):
    """Closed-form Gaussian KL (utils_bbb.py:3-5) with mu_prior = 0 and a scalar sigma_prior, as
    every reference call site uses it (linear.py:24-28, conv.py:43-47).  One fused CUDA pass that
    also produces the gradient (ops.KLFunction); takes rho (sigma = softplus(rho))."""
    if not from_rho:
        raise NotImplementedError("kl_divergence takes rho; the reference always passes softplus(self.std)")
    return ops.kl_divergence(mu, sigma_or_rho, prior_value(sigma_prior))


def prior_value(sigma_prior):
    """Host value of the (non-trainable, 1-element) `std_prior` parameter, read back ONCE per version and kept on the tensor
    object itself: a `.item()` per layer per step would stall the launch queue behind the whole forward pass."""
    if not torch.is_tensor(sigma_prior):
        return float(sigma_prior)
    cached = getattr(sigma_prior, "_qbn_host_value", None)
    if cached is None or cached[0] != (sigma_prior.data_ptr(), sigma_prior._version):
        cached = ((sigma_prior.data_ptr(), sigma_prior._version), float(sigma_prior.reshape(-1)[0]))
        sigma_prior._qbn_host_value = cached
    return cached[1]


def model_kl_divergence(model):
    """models_bbb.py:80-85,135-140,254-259: the sum of the per-layer KL terms, as one fused launch."""
    from .conv import Conv2d
    from .linear import Linear
    pairs = [(m.weight, m.std, prior_value(m.std_prior)) for m in model.modules() if isinstance(m, (Linear, Conv2d))]
    if all(p[0].is_cuda and p[0].is_contiguous() and p[1].is_contiguous() and p[0].dtype == torch.float32 for p in pairs):
        return ops.kl_divergence_multi(pairs)
    return sum(ops.kl_divergence(mu, rho, sp) for mu, rho, sp in pairs)


def softplusinv(x):
    """utils_bbb.py:7-8 (weight-sized, used once at BN-fold time: conv.py:77)."""
    return torch.log(torch.exp(x) - 1.)
