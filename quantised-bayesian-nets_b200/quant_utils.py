"""Quantisation glue of the hot path — mirror of src/quant_utils.py (reference :30-147) and the
torch.quantization pieces it instantiates, backed by libqbn kernels.

  FakeQuantize      drop-in for torch's FakeQuantize(MovingAverageMinMaxObserver, per_tensor_affine):
                    one fused observe+EMA+qparams kernel and one quantise kernel (A7), STE backward.
  QTensor           uint8 activations + (scale, zero_point) on the GPU, NHWC — what the reference keeps
                    in torch quint8 tensors on the CPU (the reference's int8 path is CPU-only).
  prepare_model / convert / postprocess_model   same call signatures as the reference.
"""
import copy

import torch
import torch.nn as nn

from . import noise, ops

UINT_BOUNDS = {8: [0, 255], 7: [0, 127], 6: [0, 63], 5: [0, 31], 4: [0, 15], 3: [0, 7], 2: [0, 3]}            # src/utils.py:18
INT_BOUNDS = {8: [-128, 127], 7: [-64, 63], 6: [-32, 31], 5: [-16, 15], 4: [-8, 7], 3: [-4, 3], 2: [-2, 1]}  # src/utils.py:19-20


class _ObserverView(nn.Module):
    """`fq.activation_post_process.min_val / max_val / eps` like torch's MovingAverageMinMaxObserver (views of the owner's
    device-side state; the state-dict keys are written by the owning FakeQuantize)."""

    def __init__(self, owner):
        super().__init__()
        object.__setattr__(self, "_owner", owner)

    @property
    def min_val(self):
        return self._owner.state[0]

    @property
    def max_val(self):
        return self._owner.state[1]

    @property
    def eps(self):
        return torch.tensor([torch.finfo(torch.float32).eps])

    def calculate_qparams(self):
        return self._owner.calculate_qparams()


class FakeQuantize(torch.ao.quantization.FakeQuantizeBase):
    """y = (clamp(rint(x/s)+z, qmin, qmax) - z) * s with an EMA(min,max; c=0.01) observer
    (torch/ao/quantization/fake_quantize.py:228-260, observer.py:374-410,668-683) in two CUDA
    kernels; gradients pass where the un-clamped integer is inside [qmin, qmax] (STE).

    A FakeQuantizeBase, so `model.apply(torch.ao.quantization.disable_observer)` and friends reach it, and its state-dict
    is torch's own key set (`fake_quant_enabled`, `observer_enabled`, `scale`, `zero_point`,
    `activation_post_process.{eps,min_val,max_val}`): a QAT checkpoint written by the reference restores the trained EMA
    range here, and the other way round."""

    def __init__(self, observer=None, quant_min=0, quant_max=255, dtype=torch.quint8, qscheme=torch.per_tensor_affine,
                 averaging_constant=0.01, **kw):
        super().__init__()
        assert qscheme in (torch.per_tensor_affine,), "the reference only uses per_tensor_affine (quant_utils.py:129-138)"
        self.quant_min, self.quant_max = int(quant_min), int(quant_max)
        self.dtype, self.qscheme = dtype, qscheme
        self.averaging_constant = float(averaging_constant)
        self.register_buffer("scale", torch.ones(1))
        self.register_buffer("zero_point", torch.zeros(1, dtype=torch.int32))
        # min, max, initialised — the observer kernel's state; saved as activation_post_process.{min_val,max_val}
        self.register_buffer("state", torch.tensor([float("inf"), float("-inf"), 0.0]), persistent=False)
        self._observer_on, self._fq_on = True, True       # host-side mirrors of the two uint8 buffers (no device sync on the hot path)
        self.register_buffer("workspace", torch.zeros(16 + 8 * 1024, dtype=torch.uint8), persistent=False)
        self.activation_post_process = _ObserverView(self)

    @classmethod
    def with_args(cls, **kwargs):
        from torch.ao.quantization.observer import _PartialWrapper
        import functools
        return _PartialWrapper(functools.partial(cls, **kwargs))

    def calculate_qparams(self):
        return self.scale.detach().clone().float(), self.zero_point.detach().clone().long()

    def enable_observer(self, enabled=True):
        self._observer_on = bool(enabled)
        self.observer_enabled[0] = 1 if enabled else 0
        return self

    def enable_fake_quant(self, enabled=True):
        self._fq_on = bool(enabled)
        self.fake_quant_enabled[0] = 1 if enabled else 0
        return self

    def disable_fake_quant(self):
        return self.enable_fake_quant(False)

    def disable_observer(self):
        return self.enable_observer(False)

    def _fq_state(self):
        st = ops.FakeQuantState.__new__(ops.FakeQuantState)
        st.qmin, st.qmax, st.c = self.quant_min, self.quant_max, self.averaging_constant
        st.state, st.scale, st.zero_point, st.workspace = self.state, self.scale, self.zero_point, self.workspace
        return st

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("qbn_b200 FakeQuantize runs on CUDA only (no CPU fallback)")
        if self.state.device != x.device:
            self.to(x.device)
        if not self._fq_on:
            if self._observer_on:         # torch observes (and refreshes scale / zero_point) even when it does not quantise
                ops.fake_quant_observe(x, self._fq_state())
            return x
        return ops.fake_quantize(x, self._fq_state(), observe=self._observer_on)

    # ---- torch's FakeQuantize / MovingAverageMinMaxObserver checkpoint keys
    def _save_to_state_dict(self, destination, prefix, keep_vars):
        super()._save_to_state_dict(destination, prefix, keep_vars)
        st = self.state.detach()
        destination[prefix + "activation_post_process.eps"] = torch.tensor([torch.finfo(torch.float32).eps])
        destination[prefix + "activation_post_process.min_val"] = st[0].clone()
        destination[prefix + "activation_post_process.max_val"] = st[1].clone()

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        lo = state_dict.pop(prefix + "activation_post_process.min_val", None)
        hi = state_dict.pop(prefix + "activation_post_process.max_val", None)
        state_dict.pop(prefix + "activation_post_process.eps", None)
        legacy = state_dict.pop(prefix + "state", None)                  # round-1 checkpoints of this package
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)
        with torch.no_grad():
            if lo is not None and hi is not None and lo.numel() == 1 and hi.numel() == 1:
                lo_f, hi_f = float(lo.reshape(-1)[0]), float(hi.reshape(-1)[0])
                seen = 1.0 if (lo_f != float("inf") and hi_f != float("-inf")) else 0.0
                self.state.copy_(torch.tensor([lo_f, hi_f, seen]))
            elif legacy is not None:
                self.state.copy_(legacy.to(self.state.device))
            elif strict and lo is None:
                missing_keys.append(prefix + "activation_post_process.min_val")
        self._observer_on = bool(int(self.observer_enabled.reshape(-1)[0]))
        self._fq_on = bool(int(self.fake_quant_enabled.reshape(-1)[0]))

    def extra_repr(self):
        return "quant_min=%d, quant_max=%d, dtype=%s" % (self.quant_min, self.quant_max, self.dtype)


# ------------------------------------------------------------------------------------------------
class QTensor:
    """Per-tensor-affine quint8 activation on the GPU: `q` uint8 (NCHW-logical, NHWC-dense, or [B,K])."""

    def __init__(self, q, scale, zero_point, bits=8):
        # `bits`: the integers are known to lie in [0, 2^bits - 1] (the producing kernel clamped them); lets
        # clamp_activation skip a full pass over the tensor when the producer already applied the model's activation width
        self.q, self.scale, self.zero_point, self.bits = q, float(scale), int(zero_point), int(bits)

    @property
    def shape(self):
        return self.q.shape

    def dim(self):
        return self.q.dim()

    def int_repr(self):
        return self.q

    def q_scale(self):
        return self.scale

    def q_zero_point(self):
        return self.zero_point

    def dequantize(self):
        return ops.dequantize_u8(self.q, self.scale, self.zero_point)

    def clamp_activation(self, args):
        """src/utils.py:25-30: integer clamp to [0, 2^a - 1] (qparams unchanged)."""
        lo, hi = UINT_BOUNDS[args.activation_precision]
        if self.bits <= args.activation_precision:          # already inside [0, 2^a - 1]: the clamp is the identity
            return self
        return QTensor(torch.clamp(self.q, lo, hi), self.scale, self.zero_point, args.activation_precision)

    def reshape(self, *shape):
        return QTensor(self.q.reshape(*shape), self.scale, self.zero_point, self.bits)

    # ---- enough of torch's quantised-tensor protocol for the REFERENCE's own model code to run on QTensors unchanged:
    # `clamp_activation` (src/utils.py:25-30: x.dtype == torch.quint8, x.q_scale(), x.q_zero_point(), torch.clamp with float
    # bounds), nn.ReLU / nn.AvgPool2d / nn.MaxPool2d between the layers (models_bbb.py:209, models_mc.py:177), Flatten.
    dtype = torch.quint8
    is_quantized = True

    @property
    def device(self):
        return self.q.device

    @property
    def is_cuda(self):
        return self.q.is_cuda

    def _clamp_float(self, lo, hi):
        """torch.clamp on a quint8 tensor with float bounds = integer clamp to the quantised bounds (qparams unchanged)."""
        def to_q(v, default):
            if v is None:
                return default
            return int(max(0, min(255, round(float(v) / self.scale) + self.zero_point)))
        lo_q, hi_q = to_q(lo, 0), to_q(hi, 255)
        bits = self.bits
        if lo_q == 0 and hi_q + 1 == (hi_q + 1 & -(hi_q + 1)):          # [0, 2^b - 1]: remember the width
            b = (hi_q + 1).bit_length() - 1
            if self.bits <= b:
                return self                                              # already inside: the clamp is the identity
            bits = b
        return QTensor(torch.clamp(self.q, lo_q, hi_q), self.scale, self.zero_point, bits)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        import torch.nn.functional as F
        x = args[0] if args else kwargs.get("input")
        if func in (torch.clamp, torch.Tensor.clamp):
            lo = args[1] if len(args) > 1 else kwargs.get("min")
            hi = args[2] if len(args) > 2 else kwargs.get("max")
            return x._clamp_float(lo, hi)
        if func in (torch.relu, F.relu, torch.Tensor.relu):
            sb = noise.sample_batch_state()
            bits = min(x.bits, sb[3]) if sb is not None else x.bits
            return QTensor(ops.i8_relu(x.q, x.zero_point, act_bits=bits), x.scale, x.zero_point, bits)
        if func in (F.avg_pool2d, torch._C._nn.avg_pool2d):
            k = args[1] if len(args) > 1 else kwargs.get("kernel_size")
            k = k if isinstance(k, int) else k[0]
            return QTensor(ops.i8_avgpool(x.q, x.zero_point, k, act_bits=x.bits), x.scale, x.zero_point, x.bits)
        if func in (F.max_pool2d, torch.max_pool2d):
            k = args[1] if len(args) > 1 else kwargs.get("kernel_size")
            st = args[2] if len(args) > 2 else kwargs.get("stride", None)
            q = F.max_pool2d(x.q.float(), k, st).to(torch.uint8).contiguous(memory_format=torch.channels_last)
            return QTensor(q, x.scale, x.zero_point, x.bits)            # order-preserving affine map: max of the integers
        if func in (torch.flatten,):
            return QTensor(torch.flatten(x.q, *args[1:], **kwargs), x.scale, x.zero_point, x.bits)
        return NotImplemented

    def size(self, i=None):
        return self.q.size() if i is None else self.q.size(i)


class Quantize(nn.Module):
    """nnq.Quantize (QuantStub after convert): float -> QTensor with the calibrated (scale, zp)."""

    def __init__(self, scale, zero_point, qmin=0, qmax=255):
        super().__init__()
        self.scale, self.zero_point, self.qmin, self.qmax = float(scale), int(zero_point), qmin, qmax

    def forward(self, x):
        xc = ops.nhwc(x.float())
        return QTensor(ops.quantize_u8(xc, self.scale, self.zero_point, self.qmin, self.qmax), self.scale, self.zero_point)

    @classmethod
    def from_float(cls, mod):
        s, z = mod.activation_post_process.calculate_qparams()
        return cls(float(s), int(z))

    @classmethod
    def from_torch(cls, mod):
        return cls(float(mod.scale.reshape(-1)[0]), int(mod.zero_point.reshape(-1)[0]))

    # nnq.Quantize keeps (scale, zero_point) as [1] buffers: same checkpoint keys
    def _save_to_state_dict(self, destination, prefix, keep_vars):
        super()._save_to_state_dict(destination, prefix, keep_vars)
        destination[prefix + 'scale'] = torch.tensor([self.scale])
        destination[prefix + 'zero_point'] = torch.tensor([self.zero_point])

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        self.scale = float(state_dict.pop(prefix + 'scale').reshape(-1)[0])
        self.zero_point = int(state_dict.pop(prefix + 'zero_point').reshape(-1)[0])
        super()._load_from_state_dict(state_dict, prefix, local_metadata, False, missing_keys, unexpected_keys, error_msgs)


class DeQuantize(nn.Module):
    def forward(self, x):
        return x.dequantize() if isinstance(x, QTensor) else x

    @classmethod
    def from_float(cls, mod):
        return cls()


# ------------------------------------------------------------------------------------------------
class QFunctional(nn.Module):
    """nnq.QFunctional as the reference uses it (src/utils.py:49-55, the residual add of BasicBlock): quantized::add of two
    quint8 activations with the output at this module's calibrated (scale, zero_point)."""

    def __init__(self, scale=1.0, zero_point=0):
        super().__init__()
        self.scale, self.zero_point = float(scale), int(zero_point)
        self.activation_post_process = nn.Identity()

    def add_relu(self, x, y):
        """quantized::add_relu (torch's QFunctional.add_relu): the ReLU rides the add's output clamp."""
        return self.add(x, y, relu=True)

    def add(self, x, y, relu=False):
        assert isinstance(x, QTensor) and isinstance(y, QTensor), "QFunctional.add takes two QTensor activations"
        assert x.q.shape == y.q.shape, "residual add: operand shapes differ"
        a = x.q.contiguous(memory_format=torch.channels_last) if x.q.dim() == 4 else x.q.contiguous()
        b = y.q.contiguous(memory_format=torch.channels_last) if y.q.dim() == 4 else y.q.contiguous()
        sb = noise.sample_batch_state()
        bits = sb[3] if sb is not None else 8               # MC engine: clamp to the model's activation width in the same pass
        q = ops.i8_add(a, x.scale, x.zero_point, b, y.scale, y.zero_point, self.scale, self.zero_point, act_bits=bits, relu=relu)
        return QTensor(q, self.scale, self.zero_point, bits)

    def mul_scalar(self, x, y):
        """quantized::mul_scalar with a positive scalar (dropout.py:39): the integers stay, the scale is multiplied (fp32)."""
        y = float(y.detach().reshape(-1)[0]) if torch.is_tensor(y) else float(y)
        if y <= 0:
            raise NotImplementedError("QFunctional.mul_scalar: positive scalars only (the reference multiplies by 1/(1-p))")
        return QTensor(x.q, x.scale * y, x.zero_point, x.bits)                # ATen: the quantizer's scale (double) times the scalar

    def extra_repr(self):
        return "scale={}, zero_point={}".format(self.scale, self.zero_point)

    @classmethod
    def from_float(cls, mod):
        s, z = mod.activation_post_process.calculate_qparams()
        return cls(float(s), int(z))

    @classmethod
    def from_torch(cls, mod):
        """torch's own nnq.QFunctional (a model converted by torch.quantization.convert on the CPU)."""
        return cls(float(mod.scale), int(mod.zero_point))

    # checkpoint keys of nnq.QFunctional: `<prefix>scale`, `<prefix>zero_point` as 0-d tensors
    def _save_to_state_dict(self, destination, prefix, keep_vars):
        super()._save_to_state_dict(destination, prefix, keep_vars)
        destination[prefix + 'scale'] = torch.tensor(self.scale)
        destination[prefix + 'zero_point'] = torch.tensor(self.zero_point)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        self.scale = float(state_dict.pop(prefix + 'scale'))
        self.zero_point = int(state_dict.pop(prefix + 'zero_point'))
        super()._load_from_state_dict(state_dict, prefix, local_metadata, False, missing_keys, unexpected_keys, error_msgs)


def qconfig_for(args):
    """quant_utils.py:129-138: activations quint8 [0, 2^a-1], weights qint8 [-2^(w-1), 2^(w-1)-1]."""
    assert 2 <= args.activation_precision <= 7 and 2 <= args.weight_precision <= 8     # quant_utils.py:120-121
    a, w = UINT_BOUNDS[args.activation_precision], INT_BOUNDS[args.weight_precision]
    return torch.ao.quantization.QConfig(
        activation=FakeQuantize.with_args(quant_min=a[0], quant_max=a[1], dtype=torch.quint8, qscheme=torch.per_tensor_affine),
        weight=FakeQuantize.with_args(quant_min=w[0], quant_max=w[1], dtype=torch.qint8, qscheme=torch.per_tensor_affine))


def _mappings():
    from .stochastic.bbb import conv as C, linear as L
    from .stochastic.bbb.quantized import conv_q, conv_qat, linear_q, linear_qat
    qat = {L.Linear: linear_qat.Linear, L.LinearReLU: linear_qat.LinearReLU, C.Conv2d: conv_qat.Conv2d, C.ConvReLU2d: conv_qat.ConvReLU2d,
           C.ConvBn2d: conv_qat.ConvBn2d, C.ConvBnReLU2d: conv_qat.ConvBnReLU2d}                      # quant_utils.py:38-43
    static = {linear_qat.Linear: linear_q.Linear, linear_qat.LinearReLU: linear_q.LinearReLU, conv_qat.Conv2d: conv_q.Conv2d,
              conv_qat.ConvReLU2d: conv_q.ConvReLU2d, conv_qat.ConvBn2d: conv_q.Conv2d, conv_qat.ConvBnReLU2d: conv_q.ConvReLU2d,
              torch.ao.quantization.QuantStub: Quantize, torch.ao.quantization.DeQuantStub: DeQuantize}   # quant_utils.py:45-54
    return qat, static


def convert(model, mapping=None, inplace=True):
    """quant_utils.py:62-99: recursive swap of every module whose type is in `mapping` via from_float."""
    if mapping is None:
        mapping = _mappings()[1]
    if not inplace:
        model = copy.deepcopy(model)
    for name, child in list(model.named_children()):
        if type(child) in mapping:
            new = mapping[type(child)].from_float(child)
            model._modules[name] = new
        elif isinstance(child, torch.ao.nn.quantized.FloatFunctional) and hasattr(child.activation_post_process, "calculate_qparams"):
            model._modules[name] = QFunctional.from_float(child)       # torch's default mapping FloatFunctional -> QFunctional
        else:
            convert(child, mapping, inplace=True)
    return model


def prepare_model(model, args, q=None, at=None):
    """quant_utils.py:112-147 for the BBB families: fuse, attach the QConfig, give every quantisable
    leaf an activation fake-quant, and swap the float BBB modules for their QAT versions."""
    if hasattr(model, "fuse_model"):
        model.fuse_model()
    qconfig = qconfig_for(args)
    model.qconfig = qconfig
    qat_map, _ = _mappings()

    def attach(mod):
        for name, child in list(mod.named_children()):
            if type(child) in qat_map:
                child.qconfig = qconfig
                # observed like torch.quantization.prepare does for leaf modules in the allow-list
                target = child[-1] if isinstance(child, nn.Sequential) and type(child[-1]) is nn.ReLU else child
                mod._modules[name] = qat_map[type(child)].from_float(_with_observer(child, qconfig))
            elif isinstance(child, torch.ao.quantization.QuantStub):
                child.qconfig = qconfig
                child.add_module("activation_post_process", qconfig.activation())
                child.register_forward_hook(lambda m, i, o: m.activation_post_process(o))
            elif type(child).__name__ == "Add" and hasattr(child, "add"):
                child.add.activation_post_process = qconfig.activation()
            else:
                attach(child)
    attach(model)
    return model


def _with_observer(mod, qconfig):
    """torch.quantization.prepare leaves `activation_post_process` on the module (or on the trailing
    ReLU of a fused Sequential) — the from_float methods of the QAT classes read it from there."""
    inner = mod[0] if isinstance(mod, nn.Sequential) else mod
    inner.qconfig = qconfig
    if not hasattr(inner, "activation_post_process"):
        inner.activation_post_process = qconfig.activation()
    mod.qconfig = qconfig
    return mod


def load_model(model, model_path, replace=True):
    """src/utils.py:112-123: load a checkpoint written by the reference (`weights*.pt`, float or converted int8) into
    `model`, keeping only the keys the model has; `model_path` may also be an already loaded state-dict.  For an int8
    checkpoint build the skeleton the way the reference does — prepare_model(model, args); model.cuda(); convert(model) —
    then call this; int8 tensors and every quantisation parameter come from the file."""
    state_dict = model_path if isinstance(model_path, dict) else torch.load(model_path, map_location=torch.device('cpu'))
    own = model.state_dict()
    for key, value in state_dict.items():
        if replace:
            key = key.replace('module.', '').replace('main_net.', '')
        if key in own:
            own[key] = value
    model.load_state_dict(own)
    return model


def postprocess_model(model, args, q=None, at=None, special_info=""):
    """quant_utils.py:101-110 without the file I/O: convert the (trained, calibrated) QAT model to int8."""
    return convert(model)


def to_device_int8(model, device="cuda"):
    """A model converted by torch itself — the reference's MC-Dropout and SGHMC families go through stock
    `torch.quantization.prepare_qat` / `convert` (src/quant_utils.py:140-141, models_mc.py:222, models_sgld.py:210) and run on
    FBGEMM on the CPU — re-housed on the GPU kernels, in place: nnq.Conv2d / nniq.ConvReLU2d / nnq.Linear / nniq.LinearReLU ->
    stochastic.quantized_det (same integers, same requantisation), nnq.Quantize / DeQuantize / QFunctional -> the classes above,
    the reference's BernoulliDropout -> the drop-in (its converted mul_mask qparams kept).  The model's own forward (the
    reference's code, `clamp_activation` included) then runs unchanged on QTensor activations."""
    import torch.ao.nn.intrinsic.quantized as nniq
    import torch.ao.nn.quantized as nnq
    from .stochastic import quantized_det as det
    from .stochastic.mcdropout.dropout import BernoulliDropout
    table = {nniq.ConvReLU2d: det.QuantizedConvReLU2d, nnq.Conv2d: det.QuantizedConv2d, nniq.LinearReLU: det.QuantizedLinearReLU,
             nnq.Linear: det.QuantizedLinear, nnq.Quantize: Quantize, nnq.DeQuantize: DeQuantize, nnq.QFunctional: QFunctional}

    def swap(mod):
        for name, child in list(mod.named_children()):
            cls = table.get(type(child))
            if cls is not None:
                new = cls.from_torch(child) if cls is not DeQuantize else DeQuantize()
                mod._modules[name] = new.to(device) if isinstance(new, nn.Module) else new
            elif type(child).__name__ == "BernoulliDropout" and not isinstance(child, BernoulliDropout):
                new = BernoulliDropout(float(child.p.detach().reshape(-1)[0]))
                for fn in ("mul_mask", "mul_scalar"):
                    f = getattr(child, fn)
                    setattr(new, fn, QFunctional.from_torch(f) if isinstance(f, nnq.QFunctional) else f)
                mod._modules[name] = new.to(device)
            else:
                if isinstance(child, BernoulliDropout):
                    for fn in ("mul_mask", "mul_scalar"):
                        f = getattr(child, fn)
                        if isinstance(f, nnq.QFunctional):
                            setattr(child, fn, QFunctional.from_torch(f))
                swap(child)
    swap(model)
    for p_ in model.parameters():
        p_.data = p_.data.to(device)
    for b_ in model.buffers():
        b_.data = b_.data.to(device)
    return model
