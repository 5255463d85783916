"""Drop-in for src/models/stochastic/sgld/utils_sgld.py: the SGHMC optimiser of the reference's SGLD family — same constructor,
same `step(burn_in=, resample_momentum=, resample_prior=)` (src/trainer.py:119-121), same per-parameter state names
(`tau`, `g`, `V_hat`, `v_momentum`, `weight_decay`, `iteration`) — with the ~35 elementwise launches per parameter tensor of
the reference fused into ONE kernel (`qbn_sghmc_step`); the Gaussian draws come from the device Philox stream keyed by
(seed, parameter index, step), or from `noise.inject` when a test pins them."""
import ctypes

import torch
from numpy.random import gamma
from torch.optim import Optimizer

from ... import _lib, noise


class SGLD(Optimizer):
    def __init__(self, params, lr=1e-2, base_C=0.05, gauss_sig=0.1, alpha0=10, beta0=10):
        self.eps = 1e-6
        self.alpha0, self.beta0 = alpha0, beta0
        self.weight_decay = 0 if gauss_sig == 0 else 1 / (gauss_sig ** 2)
        if self.weight_decay <= 0.0:
            raise ValueError("Invalid weight_decay value: {}".format(self.weight_decay))
        if lr < 0.0:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if base_C < 0:
            raise ValueError("Invalid friction term: {}".format(base_C))
        super().__init__(params, dict(lr=lr, base_C=base_C))
        self._step_no = 0

    @torch.no_grad()
    def step(self, burn_in=False, resample_momentum=False, resample_prior=False):
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
        index = 0
        for group in self.param_groups:
            for p in group["params"]:
                index += 1
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise _lib.QbnError("qbn_b200 SGLD runs on CUDA parameters only (no CPU fallback)")
                state = self.state[p]
                if len(state) == 0:
                    state["iteration"] = 0
                    state["tau"], state["g"], state["V_hat"] = torch.ones_like(p), torch.ones_like(p), torch.ones_like(p)
                    state["v_momentum"] = torch.zeros_like(p)
                    state["weight_decay"] = self.weight_decay
                if resample_prior:                      # utils_sgld.py:48-53 (host-side Gamma draw, every resample_prior_iterations)
                    alpha = self.alpha0 + p.data.nelement() / 2
                    beta = self.beta0 + (p.data ** 2).sum().item() / 2
                    state["weight_decay"] = gamma(shape=alpha, scale=1 / (beta + self.eps), size=None)
                if not (p.is_contiguous() and p.grad.is_contiguous() and p.dtype == torch.float32):
                    raise _lib.QbnError("SGLD: contiguous fp32 parameters and gradients only")
                z_m = noise.pop_injected() if (resample_momentum and noise._state["queue"] is not None) else None
                z_n = noise.pop_injected() if noise._state["queue"] is not None else None
                _lib.call("qbn_sghmc_step", ptr(p), ptr(p.grad), ptr(state["tau"]), ptr(state["g"]), ptr(state["V_hat"]), ptr(state["v_momentum"]),
                          p.numel(), float(state["weight_decay"]), float(group["lr"]), float(group["base_C"]), float(self.eps), int(bool(burn_in)),
                          int(bool(resample_momentum)), ptr(z_m.contiguous().float() if z_m is not None else None),
                          ptr(z_n.contiguous().float() if z_n is not None else None), noise.seed(), 0x5C1D0000 + index, 2 * self._step_no, stream)
        self._step_no += 1
        return None
