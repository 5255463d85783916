"""Drop-in for the metric containers of src/metrics.py:262-504, backed by device-side accumulators.

The reference keeps one torchmetrics object per metric (five tiny kernel chains + a one_hot allocation per batch,
src/metrics.py:20-29,48-57,76-85,104-112,381-383) and three host-side AverageMeters that call `.item()` on the objective,
its data term and the KL every step (src/metrics.py:351 — a device sync per step).  Here ONE kernel per batch
(`qbn_cls_metrics` / `qbn_reg_metrics`) adds every (sum, count) pair into a single device buffer, the objective meters
accumulate on the device too, and nothing is read back until somebody asks for a value.

The call sites of the reference work unchanged:
  * `src/trainer.py:128,131`            metric.update(output=, target=, obj=, kl=, main_obj=)
  * `src/trainer.py:57-70`              reset(), get_str(), scalar_logging(info, epoch), get_key_metric()
  * `experiments/utils.py:372-375`      metric.error.compute().item(), metric.ece / .entropy / .nll / .rmse
Extensions (not in the reference): `update(..., scale=)` multiplies the probabilities first (1/S after an all-reduce of
probability SUMS), `all_reduce()` sums the states over ranks, `compute()` returns every metric as a dict of floats."""
import torch

from . import ops

METRICS_MAPPING = {
    "error": "Error [$\\downarrow$,\\%]",
    "ece": "Expected Calibration Error [$\\downarrow$,\\%]",
    "entropy": "Entropy [nats]",
    "brier": "Brier Score [$\\downarrow$]",
    "nll": "Negative LL [$\\downarrow$,nats]",
    "mse": "Mean Squared Error [$\\downarrow$]",
    "rmse": "Root Mean Squared Error [$\\downarrow$]",
    "mae": "Mean Absolute Error [$\\downarrow$]",
    "obj": "Objective [$\\downarrow$]",
    "main_obj": "Main Objective [$\\downarrow$]",
    "kl": "KL Divergence [$\\downarrow$]",
}


class AverageMeter:
    """src/metrics.py:506-520 with the same attributes (`avg`, `sum`, `cnt`), but tensor values are summed where they
    live: `update` never synchronises; `avg` / `sum` convert to float only when read."""

    def __init__(self):
        self.reset()

    def reset(self):
        self._sum, self.cnt = 0.0, 0.0

    def update(self, val, n=1):
        if torch.is_tensor(val):
            val = val.detach().reshape(-1)[0].float()       # the reference keeps val.item(): first (only) element
        self._sum = self._sum + val * n
        self.cnt += n

    @property
    def sum(self):
        return float(self._sum)

    @property
    def avg(self):
        return float(self._sum) / self.cnt if self.cnt else 0.0


class _StateMetric:
    """One metric of a container: the object `metric.error` / `.nll` / ... of the reference (a torchmetrics.Metric there).
    compute() returns a 0-d tensor like torchmetrics, so `.compute().item()` (experiments/utils.py:372) works."""

    def __init__(self, owner, name):
        self._owner, self._name = owner, name

    @property
    def device(self):
        return self._owner.device

    def to(self, device):
        self._owner.to(device)
        return self

    def reset(self):
        self._owner._reset_slots(self._name)

    def update(self, preds, target):
        self._owner._update_only(self._name, preds, target)

    def compute(self):
        return self._owner._value(self._name)

    def __call__(self, preds, target):
        self.update(preds, target)
        return self.compute()


class Metric:
    """src/metrics.py:262-353: objective / data term / KL meters + logging helpers."""

    metric_labels = ["obj", "main_obj", "kl"]

    def __init__(self, output_size, writer=None):
        self.writer = writer
        self.output_size = output_size
        self.obj, self.main_obj, self.kl = AverageMeter(), AverageMeter(), AverageMeter()
        self.metrics = [self.obj, self.main_obj, self.kl]

    def reset(self):
        for m in self.metrics:
            m.reset()

    def get_metric_value(self, metric):
        if hasattr(metric, "avg"):
            val = metric.avg
        elif hasattr(metric, "compute"):
            val = metric.compute()
        else:
            val = metric()
        return val if isinstance(val, float) else val.item()

    def scalar_logging(self, info, iteration):
        if self.writer is None:
            return
        for label, metric in zip(self.metric_labels, self.metrics):
            self.writer.add_scalar(info + "/" + METRICS_MAPPING[label], self.get_metric_value(metric), iteration)

    def get_str(self):
        return "".join("%s: %s " % (METRICS_MAPPING[label], str(self.get_metric_value(m))) for label, m in zip(self.metric_labels, self.metrics))

    def get_packed(self):
        return {label.lower(): self.get_metric_value(m) for label, m in zip(self.metric_labels, self.metrics)}

    def get_key_metric(self):
        raise NotImplementedError("This method should be implemented in the child class.")

    def update(self, obj=0.0, main_obj=0.0, kl=0.0):
        for val, meter in ((obj, self.obj), (main_obj, self.main_obj), (kl, self.kl)):
            if val is not None:
                meter.update(val, 1)


class _DeviceStateMixin:
    """(sum, count) states of all metrics of a container in ONE fp32 device buffer + a host-side sample count per slot group."""

    def _init_state(self, n_slots, device):
        self._n_slots = n_slots
        self.state = None if device is None else torch.zeros(n_slots, dtype=torch.float32, device=device)
        self._counts = {}

    @property
    def device(self):
        return self.state.device if self.state is not None else torch.device("cpu")

    def to(self, device):
        device = torch.device(device)
        if self.state is None:
            self.state = torch.zeros(self._n_slots, dtype=torch.float32, device=device)
        elif self.state.device != device:
            self.state = self.state.to(device)
        return self

    def _ensure(self, like):
        if not like.is_cuda:
            raise RuntimeError("qbn_b200.metrics: the metric reductions are CUDA kernels (no CPU fallback); got a %s tensor" % like.device)
        if self.state is None or self.state.device != like.device:
            self.to(like.device)

    def _count(self, name):
        return self._counts.get(self._slots[name][0], 0)         # keyed by slot: metrics sharing a sum (mse / rmse) share its count

    def _add_count(self, names, n):
        for key in {self._slots[name][0] for name in names}:
            self._counts[key] = self._counts.get(key, 0) + int(n)

    def all_reduce(self):
        """(sum, count) states are sums over samples (`dist_reduce_fx="sum"`, src/metrics.py:17-18): one all-reduce of the
        state buffer with the per-metric counts appended."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1) or self.state is None:
            return
        keys = sorted({idx[0] for idx in self._slots.values()})
        cnt = torch.tensor([float(self._counts.get(k, 0)) for k in keys], device=self.state.device, dtype=torch.float64)
        dist.all_reduce(self.state)
        dist.all_reduce(cnt)
        self._counts = {k: int(c) for k, c in zip(keys, cnt.tolist())}


class ClassificationMetric(_DeviceStateMixin, Metric):
    """src/metrics.py:355-426.  State layout = `qbn_cls_metrics` (include/qbn.h): [error, nll, brier, entropy sums] +
    n_bins x (sum confidence, sum correct, count) for the ECE (10 equal-width bins on the max probability, l1:
    torchmetrics.CalibrationError(n_bins=10, norm="l1"), src/metrics.py:381-383)."""

    metric_labels = ["obj", "main_obj", "kl", "nll", "error", "entropy", "brier", "ece"]
    _slots = {"error": [0], "nll": [1], "brier": [2], "entropy": [3], "ece": None}

    def __init__(self, output_size, writer=None, n_bins=10, device=None):
        Metric.__init__(self, output_size, writer)
        self.n_bins = n_bins
        self._slots = dict(self._slots, ece=list(range(4, 4 + 3 * n_bins)))
        self._init_state(4 + 3 * n_bins, device)
        self.entropy, self.ece = _StateMetric(self, "entropy"), _StateMetric(self, "ece")
        self.nll, self.brier, self.error = _StateMetric(self, "nll"), _StateMetric(self, "brier"), _StateMetric(self, "error")
        self.metrics += [self.nll, self.error, self.entropy, self.brier, self.ece]

    @property
    def count(self):
        return self._count("error")

    def reset(self):
        Metric.reset(self)
        if self.state is not None:
            self.state.zero_()
        self._counts = {}

    def _reset_slots(self, name):
        if self.state is not None:
            self.state[self._slots[name]] = 0.0
        self._counts.pop(self._slots[name][0], None)

    @torch.no_grad()
    def update(self, output, target, scale=1.0, **kwargs):
        """output [B, K] class probabilities (p-bar of the MC loop), target [B] int64."""
        Metric.update(self, **kwargs)
        output = output.detach()
        self._ensure(output)
        ops.cls_metrics_accumulate(output, target, self.state, scale, self.n_bins)
        self._add_count(self._slots, target.numel())

    @torch.no_grad()
    def _update_only(self, name, preds, target):
        self._ensure(preds)
        tmp = torch.zeros_like(self.state)
        ops.cls_metrics_accumulate(preds.detach(), target, tmp, 1.0, self.n_bins)
        idx = self._slots[name]
        self.state[idx] += tmp[idx]
        self._add_count([name], target.numel())

    def _value(self, name):
        if self.state is None:
            return torch.tensor(float("nan"))
        n = max(self._count(name), 1)
        if name != "ece":
            return self.state[self._slots[name][0]] / n
        bins = self.state[4:].reshape(self.n_bins, 3)
        # sum_b |acc_b - conf_b| * n_b / N  ==  sum_b |sum correct - sum conf| / N
        return (bins[:, 1] - bins[:, 0]).abs().sum() / n

    def get_key_metric(self):
        return self.error.compute()

    def compute(self):
        """Every metric as floats (one device->host copy of the state)."""
        if self.state is None:
            return {}
        s = self.state.double().cpu()
        out = {}
        for name in ("error", "nll", "brier", "entropy"):
            out[name] = s[self._slots[name][0]].item() / max(self._count(name), 1)
        bins = s[4:].reshape(self.n_bins, 3)
        out["ece"] = float((bins[:, 1] - bins[:, 0]).abs().sum()) / max(self._count("ece"), 1)
        return out


class RegressionMetric(_DeviceStateMixin, Metric):
    """src/metrics.py:429-504.  State = `qbn_reg_metrics`: [gaussian nll, squared error, absolute error] sums."""

    metric_labels = ["obj", "main_obj", "kl", "nll", "rmse", "mse", "mae"]
    _slots = {"nll": [0], "mse": [1], "rmse": [1], "mae": [2]}

    def __init__(self, output_size=1, writer=None, device=None):
        Metric.__init__(self, output_size, writer)
        self._init_state(3, device)
        self.rmse, self.mse = _StateMetric(self, "rmse"), _StateMetric(self, "mse")
        self.mae, self.nll = _StateMetric(self, "mae"), _StateMetric(self, "nll")
        self.metrics += [self.nll, self.rmse, self.mse, self.mae]

    @property
    def count(self):
        return self._count("nll")

    def reset(self):
        Metric.reset(self)
        if self.state is not None:
            self.state.zero_()
        self._counts = {}

    def _reset_slots(self, name):
        if self.state is not None:
            self.state[self._slots[name]] = 0.0
        self._counts.pop(self._slots[name][0], None)      # mse and rmse share a slot

    @staticmethod
    def _split(output):
        if isinstance(output, (tuple, list)):
            return output[0].detach(), output[1].detach()
        if output.dim() == 2 and output.shape[1] > 1:      # the stacked [B, 2] form the reference's sub-metrics see (metrics.py:482)
            return output[:, 0].detach(), output[:, 1].detach()
        mean = output.detach().reshape(-1)
        return mean, torch.ones_like(mean)

    @torch.no_grad()
    def update(self, output, target, **kwargs):
        """output = (mean, variance), each [B]."""
        Metric.update(self, **kwargs)
        mean, var = self._split(output)
        self._ensure(mean)
        ops.reg_metrics_accumulate(mean, var, target, self.state)
        self._add_count(self._slots, target.numel())

    @torch.no_grad()
    def _update_only(self, name, preds, target):
        mean, var = self._split(preds)
        self._ensure(mean)
        tmp = torch.zeros_like(self.state)
        ops.reg_metrics_accumulate(mean, var, target, tmp)
        idx = self._slots[name]
        self.state[idx] += tmp[idx]
        self._add_count([name], target.numel())

    def _value(self, name):
        if self.state is None:
            return torch.tensor(float("nan"))
        v = self.state[self._slots[name][0]] / max(self._count(name), 1)
        return torch.sqrt(v) if name == "rmse" else v

    def get_key_metric(self):
        return self.rmse.compute()

    def compute(self):
        if self.state is None:
            return {}
        s = self.state.double().cpu()
        n = max(self.count, 1)
        return {"nll": s[0].item() / n, "mse": s[1].item() / n, "rmse": (s[1].item() / n) ** 0.5, "mae": s[2].item() / n}
