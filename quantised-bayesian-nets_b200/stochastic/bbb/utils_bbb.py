"""Mirror of src/models/stochastic/bbb/utils_bbb.py (reference :3-8)."""
import torch

from ... import ops


def _uniform_value(t, what):
    """Host value of a prior given the way the reference passes it: a scalar, a 1-element tensor, or a weight-shaped
    constant tensor (`torch.zeros_like(w)`, `torch.ones_like(w) * std_prior`: linear.py:24-28, conv.py:43-47)."""
    if not torch.is_tensor(t):
        return float(t)
    if t.numel() == 1:
        return prior_value(t)
    lo, hi = torch.aminmax(t.detach())
    lo, hi = float(lo), float(hi)
    if lo != hi:
        raise NotImplementedError("kl_divergence: a non-uniform %s tensor is outside the hot path (every reference call site passes a constant)" % what)
    return lo


def kl_divergence(mu, sigma, mu_prior, sigma_prior):
    """utils_bbb.py:3-5, same arguments and meaning: `sigma` IS the standard deviation (the reference's layers pass
    softplus(self.std)), `mu_prior` / `sigma_prior` are the prior's mean and standard deviation.  One fused CUDA pass
    that also produces the gradients w.r.t. mu and sigma.  The drop-in layers call `kl_divergence_from_rho`, which
    fuses the softplus as well."""
    return ops.kl_divergence_sigma(mu, sigma, _uniform_value(mu_prior, "mu_prior"), _uniform_value(sigma_prior, "sigma_prior"))


def kl_divergence_from_rho(mu, rho, sigma_prior):
    """KL(N(mu, softplus(rho)^2) || N(0, sigma_prior^2)): what linear.py:24-28 / conv.py:43-47 evaluate, softplus included."""
    return ops.kl_divergence(mu, rho, prior_value(sigma_prior))


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
