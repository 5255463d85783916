"""TEST INFRASTRUCTURE — imports the UNMODIFIED reference from /root/reference (this container
only; the path does not exist on the GPU box) so that

  * oracle/qbn_oracle.py (the CPU restatement) can be validated against the real modules, and
  * oracle/make_golden.py can dump golden input/output vectors into tests/golden/.

Nothing under quantised-bayesian-nets_b200/ (the product) may import this file.

The reference pins torch==1.7.1 (requirements.txt:54); the installed torch is 2.11, so the
quantised modules need the small compatibility shim of SURVEY.md §8c, applied to *torch's
namespaces* before the reference is imported.  /root/reference itself is never edited.
"""
import os
import sys
import types

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))


def _default_root():
    """/root/reference in the build container; on the GPU box the verbatim copy oracle/build_ref.py made (oracle/_ref)."""
    if os.path.isdir(os.path.join("/root/reference", "src", "models", "stochastic")):
        return "/root/reference"
    return os.path.join(_HERE, "_ref")


REFERENCE_ROOT = os.environ.get("QBN_REFERENCE_ROOT") or _default_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "models", "stochastic"))


_shimmed = False


def apply_torch_compat_shim():
    """Names the reference imports from torch 1.7 that moved in torch 2.x."""
    global _shimmed
    if _shimmed:
        return
    import torch.ao.nn.quantized.modules.conv as aoq_conv
    import torch.nn.quantized.modules.conv as nq_conv
    import torch.quantization.quantization_mappings as qm

    # conv_q.py:7  `from torch.nn.quantized.modules.conv import _ConvNd`
    if not hasattr(nq_conv, "_ConvNd"):
        nq_conv._ConvNd = aoq_conv._ConvNd
    # quant_utils.py:4,30-60  module-level mapping tables + propagation list
    if not hasattr(qm, "QAT_MODULE_MAPPINGS"):
        qm.QAT_MODULE_MAPPINGS = qm.DEFAULT_QAT_MODULE_MAPPINGS
        qm.STATIC_QUANT_MODULE_MAPPINGS = qm.DEFAULT_STATIC_QUANT_MODULE_MAPPINGS
        qm.get_qconfig_propagation_list = qm.get_default_qconfig_propagation_list
        if hasattr(qm, "__all__"):
            qm.__all__ = list(qm.__all__) + [
                "QAT_MODULE_MAPPINGS", "STATIC_QUANT_MODULE_MAPPINGS", "get_qconfig_propagation_list"]
    # quant_utils.py:5,89  swap_module(mod, mapping) — torch 2.x wants a third positional arg
    import importlib
    aoqq = importlib.import_module("torch.ao.quantization.quantize")
    qq = importlib.import_module("torch.quantization.quantize")
    _orig_swap = aoqq.swap_module
    if getattr(qq.swap_module, "_qbn_shim", False) is False:
        def swap_module(mod, mapping, custom=None):
            return _orig_swap(mod, mapping, custom if custom is not None else {})
        swap_module._qbn_shim = True
        qq.swap_module = swap_module
    # models_bbb.py:96,143,186-188,249  fuse_modules(..., fuser_func=f(mod_list))
    _orig_fuse = torch.quantization.fuse_modules
    if not getattr(_orig_fuse, "_qbn_shim", False):
        from torch.ao.quantization.fuse_modules import fuse_modules_qat

        def fuse_modules(model, modules_to_fuse, inplace=False, fuser_func=None, **kw):
            if fuser_func is not None:
                def adapted(mod_list, is_qat=None, additional_fuser_method_mapping=None):
                    return fuser_func(mod_list)
                return _orig_fuse(model, modules_to_fuse, inplace=inplace, fuser_func=adapted, **kw)
            training = any(m.training for m in model.modules())
            if training:
                return fuse_modules_qat(model, modules_to_fuse, inplace=inplace, **kw)
            return _orig_fuse(model, modules_to_fuse, inplace=inplace, **kw)
        fuse_modules._qbn_shim = True
        torch.quantization.fuse_modules = fuse_modules
    _shimmed = True


def _install_torchmetrics_standin():
    """src/metrics.py imports torchmetrics (absent here).  Minimal stand-in for the ORACLE only:
    Metric base with add_state/compute/reset and CalibrationError(n_bins, norm='l1')."""
    if "torchmetrics" in sys.modules:
        return
    tm = types.ModuleType("torchmetrics")

    class Metric:
        def __init__(self, *a, **k):
            self._defaults = {}
            self.device = torch.device("cpu")

        def add_state(self, name, default, dist_reduce_fx=None):
            self._defaults[name] = default
            setattr(self, name, default.clone() if torch.is_tensor(default) else list(default))

        def reset(self):
            for k, v in self._defaults.items():
                setattr(self, k, v.clone() if torch.is_tensor(v) else list(v))

        def to(self, device):
            self.device = torch.device(device)
            for k in self._defaults:
                v = getattr(self, k)
                if torch.is_tensor(v):
                    setattr(self, k, v.to(device))
            return self

    class CalibrationError(Metric):
        # torchmetrics multiclass calibration error, norm='l1': equal-width bins on max-prob
        # confidence, sum_b |acc_b - conf_b| * n_b / N  (bins are (lo, hi], torchmetrics style)
        def __init__(self, n_bins=10, task="multiclass", norm="l1", num_classes=None):
            super().__init__()
            self.n_bins = n_bins
            self.add_state("confidences", [], None)
            self.add_state("accuracies", [], None)

        def update(self, preds, target):
            conf, pred = preds.max(dim=1)
            self.confidences.append(conf.float())
            self.accuracies.append((pred == target).float())

        def compute(self):
            conf = torch.cat(self.confidences)
            acc = torch.cat(self.accuracies)
            bounds = torch.linspace(0, 1, self.n_bins + 1, dtype=conf.dtype)
            idx = torch.bucketize(conf, bounds, right=True) - 1
            idx = idx.clamp(0, self.n_bins - 1)
            ece = torch.zeros(())
            for b in range(self.n_bins):
                m = idx == b
                if m.any():
                    ece = ece + (acc[m].mean() - conf[m].mean()).abs() * m.float().mean()
            return ece

    tm.Metric = Metric
    tm.CalibrationError = CalibrationError
    sys.modules["torchmetrics"] = tm


def import_reference():
    """Returns the reference's `src` package (after shimming torch)."""
    if not reference_available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    apply_torch_compat_shim()
    _install_torchmetrics_standin()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import src  # noqa: F401
    return src


class Args:
    """Stand-in for the argparse Namespace threaded through the reference."""

    def __init__(self, **kw):
        self.sigma_prior = 1.0
        self.activation_precision = 7
        self.weight_precision = 8
        self.p = 0.2
        self.q = False
        self.at = False
        self.model = "conv_lenet_bbb"
        self.task = "classification"
        self.samples = 20
        self.debug = False
        self.gamma = 1.0
        self.loss_multiplier = 1.0
        self.output_size = 10
        self.__dict__.update(kw)
