"""Host-side mirror of src/losses.py (the ELBO the trainer builds around the hot path, trainer.py:96-104): same factory,
class names and call signature `loss(output, target, kl, gamma, n_batches, n_points) -> (loss, data_term, kl_term)`.

Glue, not a kernel: the data term is a [B, K]-sized reduction and the KL enters as the 0-d tensor produced by
`qbn_kl_multi` (one launch for all layers, stochastic/bbb/utils_bbb.py), so the loss is a handful of scalar torch ops on
whatever device its inputs live on; its gradient w.r.t. the probabilities feeds the LRT backward kernels."""
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


class ClassificationLoss(Loss):
    """NLL of the (already soft-maxed) model output, with the reference's 1e-8 guard (losses.py:14-29)."""

    def data_term(self, output, target):
        return F.nll_loss(torch.log(output + 1e-8), target)


class RegressionLoss(Loss):
    """Heteroscedastic Gaussian NLL on output = (mean, var): mean_b sum_d ((t-m)^2/(var+1e-8) + log(var+1e-8))
    (losses.py:31-52)."""

    def data_term(self, output, target):
        mean, var = output[0], output[1]
        return torch.mean(torch.sum((target - mean) ** 2 / (var + 1e-8) + torch.log(var + 1e-8), 1), 0)


LOSS_FACTORY = {"classification": lambda args, scaling: ClassificationLoss(args, scaling),
                "regression": lambda args, scaling: RegressionLoss(args, scaling)}
