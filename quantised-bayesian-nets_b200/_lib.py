"""ctypes binding of libqbn.so (include/qbn.h).  No torch types cross the ABI: tensors are passed
as raw device pointers and sizes.  If the library is missing or a call fails, this module raises —
there is deliberately NO CPU or PyTorch fallback for any op."""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint32, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QBN_LIB_PATH") or os.path.join(HERE, "csrc", "libqbn.so")   # (override: A/B kernel tuning only)

QBN_MATH_FP32 = 0
QBN_MATH_TF32 = 1
QBN_FLAG_RELU = 1
QBN_FLAG_A_TF32_READY = 2
QBN_FLAG_OUT_ROUND_TF32 = 4
QBN_FLAG_OUT_PHASE_SPLIT = 8
QBN_FLAG_OUT_P4 = 16
QBN_FLAG_X_SHARED_STACKED = 32
QBN_FLAG_RELU_PRE = 64


class ConvDesc(Structure):
    _fields_ = [(n, c_int32) for n in ("B", "H", "W", "C", "N", "R", "S", "stride_h", "stride_w", "pad_h", "pad_w",
                                        "dil_h", "dil_w", "Ho", "Wo", "out_pad_h", "out_pad_w")]


class P4SampleJob(Structure):
    _fields_ = [("mu_b", c_void_p), ("sigma_b", c_void_p), ("eps", c_void_p), ("w", c_void_p), ("N", c_int32), ("C", c_int32),
                ("taps", c_int32), ("stride", c_int32), ("layer_id", c_uint32), ("n_stack", c_int32), ("chan_scale", c_void_p),
                ("cb_override", c_int32), ("w_sample_stride4", c_int32), ("s_off", c_int32), ("pad_", c_int32)]


class KLJob(Structure):
    _fields_ = [("mu", c_void_p), ("rho", c_void_p), ("d_mu", c_void_p), ("d_rho", c_void_p), ("n", c_int64), ("sigma_prior", c_float),
                ("pad_", c_int32)]


class ScrubJob(Structure):
    _fields_ = [("grad", c_void_p), ("n", c_int64)]


class MaskJob(Structure):
    _fields_ = [("out", c_void_p), ("elems", c_int64), ("site_id", c_uint32), ("pad_", c_int32)]


class I8Requant(Structure):
    _fields_ = [("s_x", c_float), ("s_w", c_float), ("z_w", c_int32), ("s_out", c_float), ("z_out", c_int32), ("relu", c_int32),
                ("act_max", c_int32), ("s_res", c_float), ("z_res", c_int32), ("s_add", c_float), ("z_add", c_int32), ("add_relu", c_int32)]


class I8SampleParams(Structure):
    _fields_ = [("s_mu", c_float), ("z_mu", c_int32), ("s_sigma", c_float), ("z_sigma", c_int32),
                ("s_eps", c_float), ("z_eps", c_int32), ("s_mul", c_float), ("z_mul", c_int32),
                ("s_add", c_float), ("z_add", c_int32), ("w_min", c_int32), ("w_max", c_int32), ("n_vec", c_int64)]


P = c_void_p
_SIGNATURES = {
    "qbn_last_error": (c_char_p, []),
    "qbn_version": (c_int, []),
    "qbn_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "qbn_philox_u32": (c_int, [P, c_int64, c_uint64, c_uint32, c_uint32, P]),
    "qbn_philox_normal": (c_int, [P, c_int64, c_uint64, c_uint32, c_uint32, P]),
    "qbn_philox_bernoulli": (c_int, [P, c_int64, c_float, c_uint64, c_uint32, c_uint32, P]),
    "qbn_weight_prep": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P, P, P, P, c_int, P]),
    "qbn_weight_grad_post": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, P, P, c_int, P]),
    "qbn_lrt_fwd": (c_int, [POINTER(ConvDesc), P, P, P, P, P, c_uint64, c_uint32, c_uint32, P, P, c_int, P]),
    "qbn_lrt_bwd_workspace_bytes": (c_size_t, [POINTER(ConvDesc)]),
    "qbn_lrt_bwd": (c_int, [POINTER(ConvDesc), P, P, P, P, P, P, c_uint64, c_uint32, c_uint32, P, P, P, P, P, c_size_t, c_int, P]),
    "qbn_sample_weights": (c_int, [P, P, c_int64, c_int, P, c_uint64, c_uint32, c_uint32, P, c_int, P]),
    "qbn_conv_fwd": (c_int, [POINTER(ConvDesc), c_int, c_int, P, P, c_int, P, P, P, c_int, P, c_float, P, c_int, P]),
    "qbn_dropout_fwd": (c_int, [P, c_int64, c_int64, c_int64, P, c_float, c_float, c_uint64, c_uint32, c_uint32, P, P, P]),
    "qbn_kl_fwd_bwd": (c_int, [P, P, c_int64, c_float, P, P, P, c_float, P]),
    "qbn_kl_sigma_fwd_bwd": (c_int, [P, P, c_int64, c_float, c_float, P, P, P, c_float, P]),
    "qbn_fake_quant_fwd": (c_int, [P, c_int64, P, c_float, c_int, c_int, c_int, P, P, P, P, P, P]),
    "qbn_fake_quant_bwd": (c_int, [P, P, c_int64, P, P]),
    "qbn_quantize_u8": (c_int, [P, c_int64, c_float, c_int32, c_int, c_int, P, P]),
    "qbn_quantize_s8": (c_int, [P, c_int64, c_float, c_int32, c_int, c_int, P, P]),
    "qbn_dequantize_u8": (c_int, [P, c_int64, c_float, c_int32, P, P]),
    "qbn_i8_sample_weights": (c_int, [P, P, c_int64, c_int, POINTER(I8SampleParams), P, c_uint64, c_uint32, c_uint32, P, P]),
    "qbn_i8_conv_fwd": (c_int, [POINTER(ConvDesc), c_int, c_int, P, c_float, c_int32, P, c_int, c_float, c_int32, P, c_float,
                                c_int32, c_int, c_int, c_int, P, P, c_int, P]),
    "qbn_i8_add": (c_int, [P, c_float, c_int32, P, c_float, c_int32, c_int64, c_int64, c_float, c_int32, c_int, c_int, P, P]),
    "qbn_i8_relu": (c_int, [P, c_int64, c_int32, c_int, c_int, P, P]),
    "qbn_i8_avgpool": (c_int, [P, c_int64, c_int, c_int, c_int, c_int, c_int32, c_int, c_int, P, P]),
    "qbn_i8_dropout": (c_int, [P, c_float, c_int32, c_int64, c_int64, c_int64, P, c_float, c_float, c_int32, c_uint64, c_uint32,
                               c_uint32, c_int, c_int, P, P]),
    "qbn_i8_dropout_mc": (c_int, [P, c_float, c_int32, c_int, c_int64, c_int64, c_int64, c_float, c_float, c_int32, c_uint64, c_uint32,
                                  c_uint32, c_int, c_int, P, P]),
    "qbn_i8_conv_p16_fwd": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, ctypes.c_longlong, c_int, P, c_int, P,
                                    POINTER(I8Requant), P, ctypes.c_longlong, c_int, P, ctypes.c_longlong, P, P]),
    "qbn_p16_weight_bytes": (c_int, [c_int, c_int, c_int, c_int, c_int, POINTER(ctypes.c_longlong)]),
    "qbn_i8_p16_block_weights": (c_int, [P, c_int, c_int, c_int, c_int, c_int, c_int, P, P]),
    "qbn_i8_p16_from_nhwc": (c_int, [P, c_int64, c_int, c_int, c_int, c_int, c_int32, c_int64, P, P]),
    "qbn_i8_p16_to_nhwc": (c_int, [P, c_int64, c_int, c_int, c_int, c_int32, c_int64, P, P]),
    "qbn_i8_p16_avgpool": (c_int, [P, c_int64, c_int, c_int, c_int, c_int32, c_int64, c_int, c_int, P, P]),
    "qbn_i8_p16_dropout": (c_int, [P, ctypes.c_longlong, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, P, c_float, c_int32, c_float, c_int, P,
                                   ctypes.c_longlong, POINTER(I8Requant), P, ctypes.c_longlong, P]),
    "qbn_set_sample_base": (c_int, [P]),
    "qbn_set_pdl": (c_int, [c_int]),
    "qbn_p4_set_window": (c_int, [c_int, c_int, c_int]),
    "qbn_sghmc_step": (c_int, [P, P, P, P, P, P, c_int64, c_float, c_float, c_float, c_float, c_int, c_int, P, P, c_uint64, c_uint32, c_uint32, P]),
    "qbn_softmax_accumulate": (c_int, [P, c_int, c_int, c_int, P, c_int, P]),
    "qbn_softmax_accumulate_window": (c_int, [P, c_int, c_int, c_int, c_int, c_int, P, c_int, P]),
    "qbn_mc_mean": (c_int, [P, c_int, c_int64, P, P]),
    "qbn_reg_mc_reduce": (c_int, [P, P, c_int, c_int64, P, P, P]),
    "qbn_elbo_cls": (c_int, [P, P, P, c_int, c_int, c_float, c_float, c_float, P, P, P]),
    "qbn_cls_metrics": (c_int, [P, P, c_int, c_int, c_float, c_int, P, P]),
    "qbn_reg_metrics": (c_int, [P, P, P, c_int64, P, P]),
    "qbn_maxpool2x2": (c_int, [P, c_int64, c_int, c_int, c_int, P, P]),
    "qbn_avgpool_all": (c_int, [P, c_int64, c_int, c_int, c_float, P, P]),
    "qbn_conv_s1_fwd": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P, c_int, P, P, P, c_int, P, P]),
    "qbn_nchw_to_nhwc": (c_int, [P, c_int64, c_int, c_int, P, P]),
    "qbn_p4_weight_floats": (c_int, [c_int, c_int, c_int, c_int, c_int, POINTER(ctypes.c_longlong)]),
    "qbn_p4_block_weights": (c_int, [P, c_int, c_int, c_int, c_int, c_int, c_int, P, P]),
    "qbn_sample_weights_blocked": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P, c_uint64, c_uint32, c_uint32, P, c_int, P]),
    "qbn_sample_weights_blocked_multi": (c_int, [P, c_int, c_int64, c_int, c_uint64, c_uint32, c_int, P]),
    "qbn_conv_p4_fwd": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, c_int64, P, c_int, P, P, P, c_int64, P, c_float,
                                c_int, P, c_int64, P]),
    "qbn_p4_shortcut_block_channels": (c_int, [c_int, c_int]),
    "qbn_conv_p4_shortcut_fwd": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, c_int64, P, P, c_int64, c_int, P, P, c_int, P,
                                         c_int64, P]),
    "qbn_scrub_nan_multi": (c_int, [P, c_int, P]),
    "qbn_kl_multi": (c_int, [P, c_int, c_int64, P, c_float, P]),
    "qbn_dropout_masks_multi": (c_int, [P, c_int, c_int64, c_int, c_float, c_uint64, c_uint32, P]),
    "qbn_lrt_noise": (c_int, [P, c_int64, c_uint64, c_uint32, c_uint32, P]),
    "qbn_p4_stage_input": (c_int, [P, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_int, ctypes.c_longlong, P, P, P]),
    "qbn_p4_stage_grad": (c_int, [P, P, P, c_uint64, c_uint32, c_uint32, c_int64, c_int, c_int, c_int, c_int, c_int, ctypes.c_longlong, P, P, P]),
    "qbn_lrt_stage_input": (c_int, [P, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_int, ctypes.c_longlong, P, P, P, P, P]),
    "qbn_lrt_stage_input_noise": (c_int, [P, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_int, ctypes.c_longlong, P, P, P, P, P, c_int64,
                                          c_uint64, c_uint32, c_uint32, P]),
    "qbn_lrt_stage_grad": (c_int, [P, P, P, c_uint64, c_uint32, c_uint32, c_int64, c_int, c_int, c_int, c_int, c_int, ctypes.c_longlong, P, P, P, P, P]),
    "qbn_lrt_p4_weight_prep": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_int), c_int, P,
                                       POINTER(ctypes.c_longlong), P]),
    "qbn_lrt_conv_p4_fwd": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P, ctypes.c_longlong, P, P, P, c_uint64, c_uint32,
                                    c_uint32, P, P, P]),
    "qbn_lrt_conv_p4_dgrad": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P, ctypes.c_longlong, P, P, P, P]),
    "qbn_lrt_conv_p4_dgrad_s2": (c_int, [c_int, c_int, c_int, c_int, c_int, P, P, ctypes.c_longlong, P, P, P, P]),
    "qbn_lrt_conv_p4_dgrad_phase": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_int), c_int, c_int, P, P, ctypes.c_longlong, P,
                                            P, P, P]),
    "qbn_w32_from_p4": (c_int, [P, P, c_int, ctypes.c_longlong, P, P, P]),
    "qbn_lrt_wgrad_p4": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P, ctypes.c_longlong, P, P, ctypes.c_longlong,
                                 P, P, P]),
    "qbn_avgpool_p4": (c_int, [P, c_int64, c_int, c_int64, c_int, c_float, P, P]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)
_lib = None


class QbnError(RuntimeError):
    pass


def load():
    """Load libqbn.so (built in-tree by __graft_entry__.build()).  Fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise QbnError("libqbn.so not found at %s — build it with `python -c \"import __graft_entry__ as g; g.build()\"`. "
                       "There is no CPU/PyTorch fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        if os.environ.get("QBN_LIB_PATH") and not hasattr(lib, name):
            continue             # A/B kernel tuning against an older build: entry points it lacks simply cannot be called
        fn = getattr(lib, name)  # AttributeError if the library does not export what include/qbn.h declares
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        msg = load().qbn_last_error()
        raise QbnError("%s failed with status %d: %s" % (what, status, msg.decode() if msg else "?"))


def call(name, *args):
    check(getattr(load(), name)(*args), name)
