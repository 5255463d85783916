"""Device-side classification / regression metric accumulators (src/metrics.py:8-229,355-504).

Same update(output, target) / compute semantics as the reference's ClassificationMetric /
RegressionMetric for the hot-path metrics, but the (sum, count) states live in ONE device buffer
updated by a single kernel per batch (no one_hot allocation, no .item() per step); values are read
back only by compute()."""
import torch

from . import ops


class ClassificationMetric:
    """error / nll / brier / entropy / ece (10 equal-width bins, l1) — metrics.py:355-426,381-383."""

    def __init__(self, output_size, n_bins=10, device="cuda"):
        self.output_size, self.n_bins = output_size, n_bins
        self.state = torch.zeros(4 + 3 * n_bins, dtype=torch.float32, device=device)
        self.count = 0

    def reset(self):
        self.state.zero_()
        self.count = 0

    @torch.no_grad()
    def update(self, output, target, scale=1.0, **kwargs):
        ops.cls_metrics_accumulate(output, target, self.state, scale, self.n_bins)
        self.count += int(target.numel())

    def all_reduce(self):
        """(sum,count) states were designed with dist_reduce_fx='sum' (metrics.py:17-18)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            buf = torch.cat([self.state, torch.tensor([float(self.count)], device=self.state.device)])
            dist.all_reduce(buf)
            self.state, self.count = buf[:-1].contiguous(), int(buf[-1].item())

    def compute(self):
        s = self.state.double().cpu()
        n = max(self.count, 1)
        bins = s[4:].reshape(self.n_bins, 3)
        ece = 0.0
        for conf_sum, acc_sum, cnt in bins.tolist():
            if cnt > 0:
                ece += abs(acc_sum / cnt - conf_sum / cnt) * cnt / n
        return {"error": s[0].item() / n, "nll": s[1].item() / n, "brier": s[2].item() / n, "entropy": s[3].item() / n, "ece": ece}


class RegressionMetric:
    """nll / mse / rmse / mae — metrics.py:119-229,468-504."""

    def __init__(self, output_size=1, device="cuda"):
        self.state = torch.zeros(3, dtype=torch.float32, device=device)
        self.count = 0

    def reset(self):
        self.state.zero_()
        self.count = 0

    @torch.no_grad()
    def update(self, output, target, **kwargs):
        mean, var = output
        ops.reg_metrics_accumulate(mean, var, target, self.state)
        self.count += int(target.numel())

    def compute(self):
        s = self.state.double().cpu()
        n = max(self.count, 1)
        return {"nll": s[0].item() / n, "mse": s[1].item() / n, "rmse": (s[1].item() / n) ** 0.5, "mae": s[2].item() / n}
