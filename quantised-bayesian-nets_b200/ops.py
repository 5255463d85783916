"""Thin tensor-level wrappers over the C ABI (include/qbn.h) + the autograd Functions.

PyTorch is plumbing here (device memory, streams, autograd graph); every arithmetic step of the
hot path runs in libqbn's sm_100a kernels.  There is no fallback: a non-CUDA tensor raises.
"""
import ctypes
import math

import torch

from . import _lib
from ._lib import (ConvDesc, I8SampleParams, QBN_FLAG_A_TF32_READY, QBN_FLAG_OUT_P4, QBN_FLAG_OUT_PHASE_SPLIT,  # noqa: F401
                   QBN_FLAG_RELU_PRE, QBN_FLAG_X_SHARED_STACKED, QBN_FLAG_OUT_ROUND_TF32, QBN_FLAG_RELU, QBN_MATH_FP32, QBN_MATH_TF32)

CL = torch.channels_last


def _ptr(t):
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.QbnError("libqbn ops need CUDA tensors (there is no CPU fallback)")


def _f32(t):
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t


def pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def conv_out(h, k, s, p, d):
    return (h + 2 * p - d * (k - 1) - 1) // s + 1


def make_desc(B, H, W, C, N, R, S, stride=(1, 1), padding=(0, 0), dilation=(1, 1)):
    stride, padding, dilation = pair(stride), pair(padding), pair(dilation)
    d = ConvDesc()
    d.B, d.H, d.W, d.C, d.N, d.R, d.S = B, H, W, C, N, R, S
    d.stride_h, d.stride_w = stride
    d.pad_h, d.pad_w = padding
    d.dil_h, d.dil_w = dilation
    d.Ho = conv_out(H, R, stride[0], padding[0], dilation[0])
    d.Wo = conv_out(W, S, stride[1], padding[1], dilation[1])
    return d


def nhwc(x):
    """Logical NCHW tensor whose memory is dense NHWC (no copy if it already is)."""
    return x.contiguous(memory_format=CL) if x.dim() == 4 else x.contiguous()


# ------------------------------------------------------------------------------------------------
# RNG hooks
# ------------------------------------------------------------------------------------------------
def philox_u32(n, seed, stream_a=0, stream_b=0, device="cuda"):
    out = torch.empty(n, dtype=torch.int32, device=device)
    _lib.call("qbn_philox_u32", _ptr(out), n, seed, stream_a, stream_b, _stream())
    return out


def philox_normal(n, seed, stream_a=0, stream_b=0, device="cuda"):
    out = torch.empty(n, dtype=torch.float32, device=device)
    _lib.call("qbn_philox_normal", _ptr(out), n, seed, stream_a, stream_b, _stream())
    return out


def philox_bernoulli(n, keep_prob, seed, stream_a=0, stream_b=0, device="cuda"):
    out = torch.empty(n, dtype=torch.float32, device=device)
    _lib.call("qbn_philox_bernoulli", _ptr(out), n, float(keep_prob), seed, stream_a, stream_b, _stream())
    return out


# ------------------------------------------------------------------------------------------------
# parameter packing
# ------------------------------------------------------------------------------------------------
def weight_prep(mu, second, second_is_sigma=False, chan_scale=None, want=("mu", "sigma", "sigma2"), round_tf32=False):
    """OIHW (or [N,K]) parameters -> packed OHWI operands.  Returns dict of flat [N*K] tensors."""
    _need_cuda(mu, second)
    mu, second = _f32(mu).contiguous(), _f32(second).contiguous()
    if mu.dim() == 2:
        N, C, R, S = mu.shape[0], mu.shape[1], 1, 1
    else:
        N, C, R, S = mu.shape
    out = {k: torch.empty(N * C * R * S, dtype=torch.float32, device=mu.device) for k in want}
    _lib.call("qbn_weight_prep", _ptr(mu), _ptr(second), int(second_is_sigma), N, C, R, S,
              _ptr(chan_scale.contiguous() if chan_scale is not None else None),
              _ptr(out.get("mu")), _ptr(out.get("sigma")), _ptr(out.get("sigma2")), int(round_tf32), _stream())
    return out


def weight_grad_post(dmu_p, dsig2_p, second, second_is_sigma, shape):
    if len(shape) == 2:
        N, C, R, S = shape[0], shape[1], 1, 1
    else:
        N, C, R, S = shape
    d_mu = torch.empty(shape, dtype=torch.float32, device=dmu_p.device)
    d_second = torch.empty(shape, dtype=torch.float32, device=dmu_p.device)
    _lib.call("qbn_weight_grad_post", _ptr(dmu_p), _ptr(dsig2_p), _ptr(second.contiguous()), int(second_is_sigma), N, C, R, S,
              _ptr(d_mu), _ptr(d_second), 0, _stream())
    return d_mu, d_second


# ------------------------------------------------------------------------------------------------
# A1-A3 local reparametrisation
# ------------------------------------------------------------------------------------------------
def _geom(x, wshape, stride, padding, dilation):
    if x.dim() == 2:
        B, C = x.shape
        H = W = 1
        N, R, S = wshape[0], 1, 1
    else:
        B, C, H, W = x.shape
        N, _, R, S = wshape
    return make_desc(B, H, W, C, N, R, S, stride, padding, dilation)


def _out_like(x, d, dtype=torch.float32, lead=None):
    if x.dim() == 2:
        shape = (d.B, d.N)
        if lead is not None:
            shape = (lead,) + shape
        return torch.empty(shape, dtype=dtype, device=x.device)
    if lead is not None:
        return torch.empty((lead * d.B, d.N, d.Ho, d.Wo), dtype=dtype, device=x.device, memory_format=CL)
    return torch.empty((d.B, d.N, d.Ho, d.Wo), dtype=dtype, device=x.device, memory_format=CL)


def lrt_forward(x, mu_p, sig2_p, bias, d, eps=None, key=(0, 0, 0), math_mode=QBN_MATH_FP32, want_std=True):
    """x already NHWC-dense.  eps (optional) laid out like the output (NHWC-dense)."""
    out = _out_like(x, d)
    std = _out_like(x, d) if want_std else None
    _lib.call("qbn_lrt_fwd", ctypes.byref(d), _ptr(x), _ptr(mu_p), _ptr(sig2_p), _ptr(bias), _ptr(eps),
              key[0], key[1], key[2], _ptr(out), _ptr(std), math_mode, _stream())
    return out, std


def lrt_backward(x, mu_p, sig2_p, g, std, d, eps=None, key=(0, 0, 0), need_dx=True, need_dbias=False, math_mode=QBN_MATH_FP32):
    nbytes = _lib.load().qbn_lrt_bwd_workspace_bytes(ctypes.byref(d))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    dx = torch.empty_like(x) if need_dx else None
    dmu_p = torch.empty_like(mu_p)
    dsig2_p = torch.empty_like(sig2_p)
    dbias = torch.empty(d.N, dtype=torch.float32, device=x.device) if need_dbias else None
    _lib.call("qbn_lrt_bwd", ctypes.byref(d), _ptr(x), _ptr(mu_p), _ptr(sig2_p), _ptr(g), _ptr(std), _ptr(eps),
              key[0], key[1], key[2], _ptr(dx), _ptr(dmu_p), _ptr(dsig2_p), _ptr(dbias), _ptr(ws), nbytes, math_mode, _stream())
    return dx, dmu_p, dsig2_p, dbias


class LRTFunction(torch.autograd.Function):
    """out = x*mu + sqrt(1e-8 + x^2*softplus(rho)^2) * eps + bias  (linear.py:32-40, conv.py:24-32)
    with the closed-form backward of SURVEY §8a row A3.  `second` is rho, or sigma itself when
    second_is_sigma (QAT: sigma was fake-quantised upstream, conv_qat.py:28-32)."""

    @staticmethod
    def forward(ctx, x, weight, second, bias, stride, padding, dilation, eps, key, math_mode, second_is_sigma, chan_scale):
        _need_cuda(x, weight, second)
        xc = nhwc(_f32(x))
        d = _geom(xc, weight.shape, stride, padding, dilation)
        eps_c = nhwc(_f32(eps)) if eps is not None else None
        ctx.d, ctx.key, ctx.math_mode, ctx.second_is_sigma = d, key, math_mode, second_is_sigma
        ctx.wshape = tuple(weight.shape)
        ctx.has_bias = bias is not None
        ctx.chan_scale = chan_scale
        ctx.planar = (math_mode == QBN_MATH_TF32 and chan_scale is None and xc.dim() == 4
                      and lrt_p4_eligible(d, need_dx=ctx.needs_input_grad[0]))
        if not ctx.planar and math_mode == QBN_MATH_TF32 and (d.C % 4 or 2 * _ceil(d.N, 16) > 512):
            math_mode = ctx.math_mode = QBN_MATH_FP32      # shapes the tcgen05 gather kernels do not take (config.tf32_eligible)
        if ctx.planar:
            # TF32 mode on the planar zero-copy kernels: operands staged once, every contraction of forward and backward on tcgen05
            out, std, x_w32, xsq_w32, eps_c = lrt_p4_forward(xc, weight, second, second_is_sigma, _f32(bias), d, eps_c, key, materialise_eps=True)
            ctx.save_for_backward(xc, x_w32, xsq_w32, std, eps_c, weight.detach(), second.detach())
            return out
        packed = weight_prep(weight, second, second_is_sigma, chan_scale, want=("mu", "sigma2"))   # backward needs them unrounded
        fw = packed if math_mode != QBN_MATH_TF32 else weight_prep(weight, second, second_is_sigma, chan_scale, want=("mu", "sigma2"), round_tf32=True)
        out, std = lrt_forward(xc, fw["mu"], fw["sigma2"], _f32(bias), d, eps_c, key, math_mode)
        ctx.save_for_backward(xc, packed["mu"], packed["sigma2"], std, eps_c, second.detach())
        return out

    @staticmethod
    def backward(ctx, g):
        gc = nhwc(_f32(g))
        need_dx = ctx.needs_input_grad[0]
        if ctx.planar:
            xc, x_w32, xsq_w32, std, eps_c, weight, second = ctx.saved_tensors
            dx, dmu_p, dsig2_p = lrt_p4_backward(xc, x_w32, xsq_w32, std, eps_c, weight, second, ctx.second_is_sigma, gc, ctx.d, ctx.key, need_dx)
            dbias = gc.sum(dim=(0, 2, 3)) if ctx.has_bias else None
        else:
            xc, mu_p, sig2_p, std, eps_c, second = ctx.saved_tensors
            # TF32 mode: dx of the stride-1 layers runs on tcgen05 (two launches of the forward kernel on flipped weights); the
            # weight gradients and every other shape stay on the fp32 CUDA-core kernels (the library decides per descriptor)
            dx, dmu_p, dsig2_p, dbias = lrt_backward(xc, mu_p, sig2_p, gc, std, ctx.d, eps_c, ctx.key, need_dx, ctx.has_bias,
                                                     ctx.math_mode)
        if ctx.chan_scale is not None:
            raise _lib.QbnError("LRTFunction.backward with chan_scale: fold the scale outside (QAT ConvBn2d does)")
        d_mu, d_second = weight_grad_post(dmu_p, dsig2_p, second, ctx.second_is_sigma, ctx.wshape)
        return dx, d_mu, d_second, dbias, None, None, None, None, None, None, None, None


# ---- A1-A3 on the planar zero-copy kernels (include/qbn.h "LRT training in TF32 mode"; csrc/lrt_p4.cu, umma_conv_p4.cu KIND_LRT,
# umma_wgrad_p4.cu).  The NHWC tensors of the module boundary are staged once per layer into planar-C4 maps.
def _ceil(a, b):
    return (a + b - 1) // b * b


def lrt_p4_plane_rows(rows, Wp):
    return _ceil(rows, 128) + 128 + 2 * (Wp + 1) + 8


def lrt_p4_eligible(d, need_dx=True):
    """Shapes the planar LRT kernels take (include/qbn.h): everything else stays on the gather kernels (qbn_lrt_fwd / qbn_lrt_bwd)."""
    if (d.dil_h, d.dil_w) != (1, 1) or d.N % 8 or d.N > 256 or d.R > 5 or d.S > 5:
        return False
    if need_dx and d.C % 8:
        return False
    if (d.stride_h, d.stride_w) == (1, 1):
        ok = d.R % 2 == 1 and d.S % 2 == 1 and (d.pad_h, d.pad_w) == ((d.R - 1) // 2, (d.S - 1) // 2)
        return ok and d.H + (d.R - 1) // 2 > 2 and d.W + (d.S - 1) // 2 > 2
    if (d.stride_h, d.stride_w) == (2, 2) and d.H % 2 == 0 and d.W % 2 == 0 and d.H >= 4 and d.W >= 4:
        return (d.R, d.S, d.pad_h, d.pad_w) in ((3, 3, 1, 1), (1, 1, 0, 0))
    return False


def _lrt_p4_geom(d):
    s2 = d.stride_h == 2
    bh, bw = (1, 1) if s2 else ((d.R - 1) // 2, (d.S - 1) // 2)
    return s2, bh, bw, d.Ho + bh, d.Wo + bw, _ceil(d.C, 8)


def lrt_p4_weight_prep(weight, second, second_is_sigma, d, mode, tap_list=None):
    C_pad = _ceil(d.C, 8)
    if mode == 0:
        n = 2 * p4_weight_floats(C_pad, d.N, d.R, d.S, d.stride_h)
    elif mode == 1:
        n = 2 * p4_weight_floats(d.N, d.C, d.R, d.S, 1)
    elif mode == 3:
        n = 8 * p4_weight_floats(d.N, d.C, 1, 4, 1)
    else:
        n = 2 * p4_weight_floats(d.N, d.C, 1, len(tap_list), 1)
    out = torch.empty(n, dtype=torch.float32, device=weight.device)
    taps = (ctypes.c_int * len(tap_list))(*tap_list) if tap_list else None
    got = ctypes.c_longlong(0)
    _lib.call("qbn_lrt_p4_weight_prep", _ptr(weight), _ptr(second), int(second_is_sigma), d.N, d.C, C_pad, d.R, d.S, d.stride_h, mode, taps,
              len(tap_list) if tap_list else 0, _ptr(out), ctypes.byref(got), _stream())
    if got.value != n:
        raise _lib.QbnError("qbn_lrt_p4_weight_prep wrote %d floats into a buffer of %d" % (got.value, n))
    return out


def lrt_p4_forward(xc, weight, second, second_is_sigma, bias, d, eps=None, key=(0, 0, 0), want_w32=True, materialise_eps=False):
    """xc NHWC-dense [B, C, H, W] (channels_last).  Returns out, std (NHWC) and the operands the backward needs: x, x^2 in the W32
    layout of the weight-gradient kernel (want_w32=False: the planar-C4 maps the forward itself read) [, eps when materialise_eps]."""
    s2, bh, bw, Hp, Wp, C_pad = _lrt_p4_geom(d)
    rows = (4 if s2 else 1) * d.B * Hp * Wp
    pr = lrt_p4_plane_rows(rows, Wp)
    x_p4 = torch.empty((C_pad // 4, pr, 4), dtype=torch.float32, device=xc.device)
    xsq_p4 = torch.empty_like(x_p4)
    drawn = False
    if want_w32:      # one pass over x writes both layouts (forward / input gradient: planar C4; weight gradients: W32)
        x_w32 = torch.empty(((C_pad + 31) // 32, pr, 32), dtype=torch.float32, device=xc.device)
        xsq_w32 = torch.empty_like(x_w32)
        if eps is None and materialise_eps and (d.B * d.Ho * d.Wo * d.N) % 4 == 0:
            # ... and the layer's noise tensor in the same launch (ALU work beside the staging's memory traffic)
            eps = _out_like(xc, d)
            _lib.call("qbn_lrt_stage_input_noise", _ptr(xc), d.B, d.H, d.W, d.C, C_pad, bh, bw, int(s2), pr, _ptr(x_p4), _ptr(xsq_p4), _ptr(x_w32),
                      _ptr(xsq_w32), _ptr(eps), eps.numel(), key[0], key[1], key[2], _stream())
            drawn = True
        else:
            _lib.call("qbn_lrt_stage_input", _ptr(xc), d.B, d.H, d.W, d.C, C_pad, bh, bw, int(s2), pr, _ptr(x_p4), _ptr(xsq_p4), _ptr(x_w32),
                      _ptr(xsq_w32), _stream())
    else:
        _lib.call("qbn_p4_stage_input", _ptr(xc), d.B, d.H, d.W, d.C, C_pad, bh, bw, int(s2), pr, _ptr(x_p4), _ptr(xsq_p4), _stream())
    w = lrt_p4_weight_prep(weight.detach().contiguous(), second.detach().contiguous(), second_is_sigma, d, 0)
    out, std = _out_like(xc, d), _out_like(xc, d)
    if eps is None and materialise_eps and not drawn:
        # the epilogue's Philox draw as a tensor (same values): generated at full occupancy, re-read by the backward
        eps = _out_like(xc, d)
        _lib.call("qbn_lrt_noise", _ptr(eps), eps.numel(), key[0], key[1], key[2], _stream())
    _lib.call("qbn_lrt_conv_p4_fwd", d.B, Hp, Wp, C_pad, d.N, d.R, d.S, d.stride_h, _ptr(x_p4), _ptr(xsq_p4), pr, _ptr(w), _ptr(bias), _ptr(eps),
              key[0], key[1], key[2], _ptr(out), _ptr(std), _stream())
    keep = (x_w32, xsq_w32) if want_w32 else (x_p4, xsq_p4)
    return (out, std) + keep + ((eps,) if materialise_eps else ())


def w32_from_p4(a, b=None):
    """planar C4 [C_pad/4, rows, 4] -> W32 [ceil(C_pad/32), rows, 32] (include/qbn.h: qbn_w32_from_p4); b: a second tensor of the same shape."""
    chunks, rows = a.shape[0], a.stride(0) // 4
    oa = torch.empty(((chunks + 7) // 8, rows, 32), dtype=torch.float32, device=a.device)
    ob = torch.empty_like(oa) if b is not None else None
    _lib.call("qbn_w32_from_p4", _ptr(a), _ptr(b), chunks * 4, rows, _ptr(oa), _ptr(ob), _stream())
    return (oa, ob) if b is not None else oa


def lrt_p4_backward(xc, x_w32, xsq_w32, std, eps, weight, second, second_is_sigma, gc, d, key=(0, 0, 0), need_dx=True):
    """Closed-form backward of SURVEY 8a row A3 on the planar kernels.  Returns dx (NHWC or None), dmu_p, dsig2_p (packed OHWI)."""
    s2, bh, bw, Hp, Wp, C_pad = _lrt_p4_geom(d)
    weight, second = weight.contiguous(), second.contiguous()
    pr_g = lrt_p4_plane_rows(d.B * Hp * Wp, Wp)
    g_p4 = torch.empty((d.N // 4, pr_g, 4), dtype=torch.float32, device=gc.device)
    dv_p4 = torch.empty_like(g_p4)
    g_w32 = torch.empty(((d.N + 31) // 32, pr_g, 32), dtype=torch.float32, device=gc.device)
    dv_w32 = torch.empty_like(g_w32)
    _lib.call("qbn_lrt_stage_grad", _ptr(gc), _ptr(std), _ptr(eps), key[0], key[1], key[2], d.B, d.Ho, d.Wo, d.N, bh, bw, pr_g, _ptr(g_p4),
              _ptr(dv_p4), _ptr(g_w32), _ptr(dv_w32), _stream())
    both = torch.empty((2, d.N * d.R * d.S * d.C), dtype=torch.float32, device=gc.device)     # adjacent: the kernel zeroes them with one memset
    dmu_p, dsig2_p = both[0], both[1]
    _lib.call("qbn_lrt_wgrad_p4", d.B, Hp, Wp, d.C, d.N, d.R, d.S, d.stride_h, _ptr(g_w32), _ptr(dv_w32), pr_g, _ptr(x_w32), _ptr(xsq_w32),
              x_w32.stride(0) // 32, _ptr(dmu_p), _ptr(dsig2_p), _stream())
    dx = None
    if need_dx:
        if not s2:
            w_t = lrt_p4_weight_prep(weight, second, second_is_sigma, d, 1)
            dx = torch.empty_like(xc)
            _lib.call("qbn_lrt_conv_p4_dgrad", d.B, Hp, Wp, d.N, d.C, d.R, d.S, _ptr(g_p4), _ptr(dv_p4), pr_g, _ptr(w_t), _ptr(xc), _ptr(dx), _stream())
        elif d.R == 3:
            # pixels (2i+a, 2j+b) of dx get the taps r = a+1 (mod 2), s = b+1 (mod 2); tap 0 of the filter reads g one row / column
            # further.  One launch: the phases are the kernel's 'samples', each with four (zero-padded) taps
            w_t = lrt_p4_weight_prep(weight, second, second_is_sigma, d, 3)
            dx = torch.empty_like(xc)
            _lib.call("qbn_lrt_conv_p4_dgrad_s2", d.B, Hp, Wp, d.N, d.C, _ptr(g_p4), _ptr(dv_p4), pr_g, _ptr(w_t), _ptr(xc), _ptr(dx), _stream())
        else:
            # 1x1 stride 2: only the pixels (2i, 2j) receive a gradient
            dx = torch.zeros_like(xc)
            w_t = lrt_p4_weight_prep(weight, second, second_is_sigma, d, 2, [0])
            _lib.call("qbn_lrt_conv_p4_dgrad_phase", d.B, Hp, Wp, d.N, d.C, 1, (ctypes.c_int * 1)(0), 0, 0, _ptr(g_p4), _ptr(dv_p4), pr_g, _ptr(w_t),
                      _ptr(xc), _ptr(dx), _stream())
    return dx, dmu_p, dsig2_p


# ------------------------------------------------------------------------------------------------
# A4 eval-time sampling + contraction
# ------------------------------------------------------------------------------------------------
def sample_weights(mu_p, sigma_p, n_samples=1, eps=None, seed=0, layer_id=0, sample0=0, round_tf32=False):
    n = mu_p.numel()
    w = torch.empty((n_samples, n), dtype=torch.float32, device=mu_p.device)
    _lib.call("qbn_sample_weights", _ptr(mu_p), _ptr(sigma_p), n, n_samples, _ptr(eps), seed, layer_id, sample0, _ptr(w), int(round_tf32), _stream())
    return w


def conv_forward(x, w, d, n_samples=1, x_shared=True, w_shared=False, scale=None, shift=None, residual=None, relu=False,
                 in_mask=None, in_mult=1.0, math_mode=QBN_MATH_FP32, out=None, flags=0):
    """x NHWC-dense ([B,..] if x_shared else [S*B,..]); w [S][N][K] packed.  Returns [S*B, N, Ho, Wo]
    (channels-last) or [S, B, N] for linear geometry."""
    if out is None:
        out = _out_like(x, d, lead=n_samples) if (n_samples > 1 or not x_shared) else _out_like(x, d)
        if x.dim() == 2 and out.dim() == 3 and n_samples == 1:
            out = out[0]
    _lib.call("qbn_conv_fwd", ctypes.byref(d), n_samples, int(x_shared), _ptr(x), _ptr(w), int(w_shared), _ptr(scale), _ptr(shift),
              _ptr(residual), int(bool(relu)) | int(flags), _ptr(in_mask), float(in_mult), _ptr(out), math_mode, _stream())
    return out


def pack_ohwi(t):
    """[N,C,R,S] (or [N,K]) tensor -> flat packed OHWI order (host-side plumbing for injected noise)."""
    if t.dim() == 4:
        return t.permute(0, 2, 3, 1).contiguous().reshape(-1)
    return t.contiguous().reshape(-1)


# ------------------------------------------------------------------------------------------------
# A5 KL
# ------------------------------------------------------------------------------------------------
class KLFunction(torch.autograd.Function):
    """utils_bbb.py:3-5 with mu_prior=0 and scalar sigma_prior; value and gradient in one pass."""

    @staticmethod
    def forward(ctx, mu, rho, sigma_prior):
        _need_cuda(mu, rho)
        mu_c, rho_c = _f32(mu).contiguous(), _f32(rho).contiguous()
        kl = torch.zeros((), dtype=torch.float32, device=mu.device)
        need = mu.requires_grad or rho.requires_grad
        d_mu = torch.zeros_like(mu_c) if need else None
        d_rho = torch.zeros_like(rho_c) if need else None
        _lib.call("qbn_kl_fwd_bwd", _ptr(mu_c), _ptr(rho_c), mu_c.numel(), float(sigma_prior), _ptr(kl), _ptr(d_mu), _ptr(d_rho),
                  1.0, _stream())
        if need:
            ctx.save_for_backward(d_mu, d_rho)
        return kl

    @staticmethod
    def backward(ctx, g):
        d_mu, d_rho = ctx.saved_tensors
        return g * d_mu, g * d_rho, None


def kl_divergence(mu, rho, sigma_prior):
    return KLFunction.apply(mu, rho, float(sigma_prior))


class KLSigmaFunction(torch.autograd.Function):
    """utils_bbb.py:3-5 with the reference's argument list (sigma, scalar priors); value and gradient in one pass."""

    @staticmethod
    def forward(ctx, mu, sigma, mu_prior, sigma_prior):
        _need_cuda(mu, sigma)
        mu_c, sg_c = _f32(mu).contiguous(), _f32(sigma).contiguous()
        kl = torch.zeros((), dtype=torch.float32, device=mu.device)
        need = mu.requires_grad or sigma.requires_grad
        d_mu = torch.zeros_like(mu_c) if need else None
        d_sg = torch.zeros_like(sg_c) if need else None
        _lib.call("qbn_kl_sigma_fwd_bwd", _ptr(mu_c), _ptr(sg_c), mu_c.numel(), float(mu_prior), float(sigma_prior), _ptr(kl), _ptr(d_mu),
                  _ptr(d_sg), 1.0, _stream())
        if need:
            ctx.save_for_backward(d_mu, d_sg)
        ctx.shapes = (mu.shape, sigma.shape)
        return kl

    @staticmethod
    def backward(ctx, g):
        d_mu, d_sg = ctx.saved_tensors
        return (g * d_mu).reshape(ctx.shapes[0]), (g * d_sg).reshape(ctx.shapes[1]), None, None


def kl_divergence_sigma(mu, sigma, mu_prior, sigma_prior):
    return KLSigmaFunction.apply(mu, sigma, float(mu_prior), float(sigma_prior))


class KLMultiFunction(torch.autograd.Function):
    """Sum of the closed-form KL over every Bayesian layer of a model (models_bbb.py:254-259) in ONE launch, with all
    gradients produced by the same pass into one flat buffer.  Inputs: sigma priors (host floats), then mu_0, rho_0, mu_1, ..."""

    _tables = {}

    @staticmethod
    def forward(ctx, priors, *tensors):
        from ._lib import KLJob
        _need_cuda(*tensors)
        dev = tensors[0].device
        key = (tuple((t.data_ptr(), t.numel()) for t in tensors), tuple(priors), dev.index)
        ent = KLMultiFunction._tables.get(key)
        if ent is None:
            if len(KLMultiFunction._tables) > 8:
                KLMultiFunction._tables.clear()
            for t in tensors:
                if t.dtype != torch.float32 or not t.is_contiguous():
                    raise _lib.QbnError("kl_divergence_multi needs contiguous fp32 parameters")
            sizes = [t.numel() for t in tensors]
            offs = [0]
            for n in sizes:
                offs.append(offs[-1] + (n + 3) // 4 * 4)
            flat = torch.zeros(offs[-1], dtype=torch.float32, device=dev)
            jobs = (KLJob * (len(tensors) // 2))()
            for i in range(len(tensors) // 2):
                mu, rho = tensors[2 * i], tensors[2 * i + 1]
                jobs[i] = KLJob(mu.data_ptr(), rho.data_ptr(), flat.data_ptr() + 4 * offs[2 * i], flat.data_ptr() + 4 * offs[2 * i + 1],
                                mu.numel(), float(priors[i]), 0)
            raw = torch.frombuffer(bytearray(bytes(jobs)), dtype=torch.uint8).to(dev)
            ent = (raw, flat, offs, sizes, max(sizes))
            KLMultiFunction._tables[key] = ent
        raw, flat, offs, sizes, max_n = ent
        kl = torch.zeros((), dtype=torch.float32, device=dev)
        _lib.call("qbn_kl_multi", _ptr(raw), len(tensors) // 2, max_n, _ptr(kl), 1.0, _stream())
        ctx.ent = ent
        ctx.shapes = [t.shape for t in tensors]
        return kl

    @staticmethod
    def backward(ctx, g):
        raw, flat, offs, sizes, _ = ctx.ent
        scaled = flat * g                      # one launch for every layer's gradient
        grads = [scaled[offs[i]:offs[i] + sizes[i]].view(ctx.shapes[i]) for i in range(len(sizes))]
        return (None, *grads)


def kl_divergence_multi(pairs):
    """pairs: [(mu, rho, sigma_prior float), ...] -> scalar sum of the per-layer KL terms."""
    priors = tuple(float(p[2]) for p in pairs)
    tensors = []
    for mu, rho, _ in pairs:
        tensors += [mu, rho]
    return KLMultiFunction.apply(priors, *tensors)


# ------------------------------------------------------------------------------------------------
# A8 MC-Dropout
# ------------------------------------------------------------------------------------------------
def dropout_forward(x, p, mask=None, key=(0, 0, 0), return_mask=False):
    """dropout.py:15-40 (float branch).  x NCHW-logical/NHWC-dense or [B,C]; mask [B,C] injected or Philox.
    return_mask: also return the [B,C] mask that was applied (the backward multiplies the gradient by it)."""
    _need_cuda(x)
    xc = nhwc(_f32(x))
    if xc.dim() == 4:
        B, C, H, W = xc.shape
        hw = H * W
    else:
        B, C = xc.shape
        hw = 1
    out = torch.empty_like(xc)
    mult = float((torch.ones(1) / (1.0 - torch.ones(1) * p)).item())  # dropout.py:10, fp32 like the Parameter
    mask_out = torch.empty((B, C), dtype=torch.float32, device=x.device) if mask is None else None
    _lib.call("qbn_dropout_fwd", _ptr(xc), B, hw, C, _ptr(mask.contiguous() if mask is not None else None), float(1.0 - p), mult,
              key[0], key[1], key[2], _ptr(out), _ptr(mask_out), _stream())
    if return_mask:
        return out, (mask_out if mask is None else mask)
    return out


class DropoutFunction(torch.autograd.Function):
    """y = x * mask * 1/(1-p) (dropout.py:35-39); dy/dx = mask * 1/(1-p): the same kernel with the saved mask."""

    @staticmethod
    def forward(ctx, x, p, mask, key):
        out, used = dropout_forward(x.detach(), p, mask, key, return_mask=True)
        ctx.save_for_backward(used)
        ctx.p = p
        return out

    @staticmethod
    def backward(ctx, g):
        (used,) = ctx.saved_tensors
        return dropout_forward(g, ctx.p, used), None, None, None


# ------------------------------------------------------------------------------------------------
# A7 fake quantisation
# ------------------------------------------------------------------------------------------------
class FakeQuantState:
    """Device-side MovingAverageMinMaxObserver + qparams (observer.py:374-410,668-683)."""

    def __init__(self, qmin, qmax, averaging_constant=0.01, device="cuda"):
        self.qmin, self.qmax, self.c = int(qmin), int(qmax), float(averaging_constant)
        self.state = torch.tensor([math.inf, -math.inf, 0.0], dtype=torch.float32, device=device)
        self.scale = torch.ones(1, dtype=torch.float32, device=device)
        self.zero_point = torch.zeros(1, dtype=torch.int32, device=device)
        self.workspace = torch.zeros(16 + 8 * 1024, dtype=torch.uint8, device=device)


class FakeQuantFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, fq, observe):
        _need_cuda(x)
        xc = _f32(x).contiguous()
        y = torch.empty_like(xc)
        mask = torch.empty(xc.shape, dtype=torch.uint8, device=x.device)
        _lib.call("qbn_fake_quant_fwd", _ptr(xc), xc.numel(), _ptr(fq.state), fq.c, int(observe), fq.qmin, fq.qmax,
                  _ptr(fq.scale), _ptr(fq.zero_point), _ptr(y), _ptr(mask), _ptr(fq.workspace), _stream())
        ctx.save_for_backward(mask)
        return y

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        gc = _f32(g).contiguous()
        gx = torch.empty_like(gc)
        _lib.call("qbn_fake_quant_bwd", _ptr(gc), _ptr(mask), gc.numel(), _ptr(gx), _stream())
        return gx, None, None


def fake_quantize(x, fq, observe=True):
    return FakeQuantFunction.apply(x, fq, observe)


def fake_quant_observe(x, fq):
    """Observer pass only (min/max EMA + scale/zero_point), no quantisation: FakeQuantize with fake-quant disabled."""
    _need_cuda(x)
    xc = _f32(x.detach()).contiguous()
    _lib.call("qbn_fake_quant_fwd", _ptr(xc), xc.numel(), _ptr(fq.state), fq.c, 1, fq.qmin, fq.qmax, _ptr(fq.scale), _ptr(fq.zero_point),
              None, None, _ptr(fq.workspace), _stream())


# ------------------------------------------------------------------------------------------------
# A6 int8
# ------------------------------------------------------------------------------------------------
def quantize_u8(x, scale, zp, qmin=0, qmax=255):
    xc = _f32(x).contiguous()
    q = torch.empty(xc.shape, dtype=torch.uint8, device=x.device)
    _lib.call("qbn_quantize_u8", _ptr(xc), xc.numel(), float(scale), int(zp), qmin, qmax, _ptr(q), _stream())
    return q


def quantize_s8(x, scale, zp, qmin=-128, qmax=127):
    xc = _f32(x).contiguous()
    q = torch.empty(xc.shape, dtype=torch.int8, device=x.device)
    _lib.call("qbn_quantize_s8", _ptr(xc), xc.numel(), float(scale), int(zp), qmin, qmax, _ptr(q), _stream())
    return q


def dequantize_u8(q, scale, zp):
    x = torch.empty(q.shape, dtype=torch.float32, device=q.device)
    _lib.call("qbn_dequantize_u8", _ptr(q), q.numel(), float(scale), int(zp), _ptr(x), _stream())
    return x


def i8_sample_weights(mu_q, sigma_q, params, n_samples=1, eps=None, seed=0, layer_id=0, sample0=0):
    """mu_q/sigma_q flat int8 in packed OHWI order; params: I8SampleParams."""
    n = mu_q.numel()
    w = torch.empty((n_samples, n), dtype=torch.int8, device=mu_q.device)
    _lib.call("qbn_i8_sample_weights", _ptr(mu_q), _ptr(sigma_q), n, n_samples, ctypes.byref(params), _ptr(eps), seed, layer_id,
              sample0, _ptr(w), _stream())
    return w


def i8_conv_forward(x_q, s_x, z_x, w_q, s_w, z_w, d, bias, s_out, z_out, relu, act_bits=7, n_samples=1, x_shared=True,
                    w_shared=False, want_acc=False, path=0, linear=False, x_bits=8):
    """x_q uint8 NHWC-dense; w_q int8 [S][N][K] packed.  Returns uint8 NHWC-dense output (+ int32 acc).
    x_bits: width the INPUT integers are known to fit (QTensor.bits).  The tcgen05 kind::i8 kernel stages (x - z_x) as s8,
    which only holds for 7-bit inputs; the library's AUTO choice looks at the OUTPUT clamp, so an input that may use the
    full quint8 range (Quantize output, un-clamped model) is pinned to the exact CUDA-core kernel here."""
    if path == 0 and int(x_bits) > 7:
        path = 1        # QBN_I8_IMAD
    lead = n_samples if (n_samples > 1 or not x_shared) else None
    if linear:
        shape = (d.B, d.N) if lead is None else (lead, d.B, d.N)
        out = torch.empty(shape, dtype=torch.uint8, device=x_q.device)
        acc = torch.empty(shape, dtype=torch.int32, device=x_q.device) if want_acc else None
    else:
        nb = d.B if lead is None else lead * d.B
        out = torch.empty((nb, d.N, d.Ho, d.Wo), dtype=torch.uint8, device=x_q.device, memory_format=CL)
        acc = torch.empty((nb, d.N, d.Ho, d.Wo), dtype=torch.int32, device=x_q.device, memory_format=CL) if want_acc else None
    amax = (1 << act_bits) - 1
    _lib.call("qbn_i8_conv_fwd", ctypes.byref(d), n_samples, int(x_shared), _ptr(x_q), float(s_x), int(z_x), _ptr(w_q), int(w_shared),
              float(s_w), int(z_w), _ptr(bias), float(s_out), int(z_out), int(relu), 0, amax, _ptr(out), _ptr(acc), int(path), _stream())
    return (out, acc) if want_acc else out


def i8_add(a, sa, za, b, sb, zb, so, zo, act_bits=7, n_vec=-1, relu=False):
    """quantized::add; relu=True is quantized::add_relu: on quint8 the ReLU is a floor at the output zero point."""
    out = torch.empty_like(a)
    _lib.call("qbn_i8_add", _ptr(a), float(sa), int(za), _ptr(b), float(sb), int(zb), a.numel(), n_vec, float(so), int(zo),
              max(0, int(zo)) if relu else 0, (1 << act_bits) - 1, _ptr(out), _stream())
    return out


def i8_relu(x_q, z_x, act_bits=7):
    """quint8 ReLU + clamp_activation in one pass (any layout: elementwise)."""
    out = torch.empty_like(x_q)
    _lib.call("qbn_i8_relu", _ptr(x_q), x_q.numel(), int(z_x), 0, (1 << act_bits) - 1, _ptr(out), _stream())
    return out


def i8_avgpool(x_q, z_x, k, act_bits=7):
    """x_q uint8 NCHW-logical / NHWC-dense [B,C,H,W] -> [B,C,H/k,W/k] (same memory format)."""
    assert x_q.dim() == 4 and x_q.dtype == torch.uint8
    x_q = x_q.contiguous(memory_format=torch.channels_last)
    B, C, H, W = x_q.shape
    out = torch.empty((B, C, H // k, W // k), dtype=torch.uint8, device=x_q.device, memory_format=torch.channels_last)
    _lib.call("qbn_i8_avgpool", _ptr(x_q), B, H, W, C, int(k), int(z_x), 0, (1 << act_bits) - 1, _ptr(out), _stream())
    return out


def i8_dropout(x_q, s_x, z_x, p, s_m, z_m, mask=None, key=(0, 0, 0), act_bits=7):
    if x_q.dim() == 4:
        B, C, H, W = x_q.shape
        hw = H * W
    else:
        B, C = x_q.shape
        hw = 1
    out = torch.empty_like(x_q)
    _lib.call("qbn_i8_dropout", _ptr(x_q), float(s_x), int(z_x), B, hw, C, _ptr(mask), float(1.0 - p), float(s_m), int(z_m),
              key[0], key[1], key[2], 0, (1 << act_bits) - 1, _ptr(out), _stream())
    return out


def i8_dropout_batched(x_q, s_x, z_x, p, s_m, z_m, n_samples, key, act_bits=8):
    """int8 MC-Dropout of a chunk of samples: x_q [n_samples*B, C(, H, W)], key = (seed, site, sample0)."""
    if x_q.dim() == 4:
        rows, C, H, W = x_q.shape
        hw = H * W
    else:
        rows, C = x_q.shape
        hw = 1
    assert rows % n_samples == 0
    out = torch.empty_like(x_q)
    _lib.call("qbn_i8_dropout_mc", _ptr(x_q), float(s_x), int(z_x), int(n_samples), rows // n_samples, hw, C, float(1.0 - p), float(s_m),
              int(z_m), key[0], key[1], key[2], 0, (1 << act_bits) - 1, _ptr(out), _stream())
    return out


# ------------------------------------------------------------------------------------------------
# A9 / A10
# ------------------------------------------------------------------------------------------------
def softmax_accumulate(logits, psum=None, window=None):
    """logits [S,B,K] -> psum [B,K] (+)= sum_s softmax(logits_s).  window = (first_img, end_img): sample 0 contributes images
    [first_img, B) only and sample S-1 images [0, end_img) only (the unit window of dist.shard_units)."""
    S, B, K = logits.shape
    acc = psum is not None
    if psum is None:
        psum = torch.empty((B, K), dtype=torch.float32, device=logits.device)
    first, end = window if window is not None else (0, B)
    _lib.call("qbn_softmax_accumulate_window", _ptr(logits.contiguous()), S, B, K, int(first), int(end), _ptr(psum), int(acc), _stream())
    return psum


def mc_mean(probs):
    """probs [S, ...] -> mean over S (experiments/utils.py:355)."""
    S = probs.shape[0]
    pc = _f32(probs).contiguous()
    out = torch.empty(pc.shape[1:], dtype=torch.float32, device=probs.device)
    _lib.call("qbn_mc_mean", _ptr(pc), S, out.numel(), _ptr(out), _stream())
    return out


def reg_mc_reduce(mu, var):
    """mu, var [S,B] -> (mean, var_total) (experiments/utils.py:349-353)."""
    S = mu.shape[0]
    muc, varc = _f32(mu).contiguous(), _f32(var).contiguous()
    n = muc.numel() // S
    mean = torch.empty(muc.shape[1:], dtype=torch.float32, device=mu.device)
    vout = torch.empty(muc.shape[1:], dtype=torch.float32, device=mu.device)
    _lib.call("qbn_reg_mc_reduce", _ptr(muc), _ptr(varc), S, n, _ptr(mean), _ptr(vout), _stream())
    return mean, vout


def cls_metrics_accumulate(probs, target, out, scale=1.0, n_bins=10):
    """out [4+3*n_bins] fp32 device accumulator (see include/qbn.h)."""
    B, K = probs.shape
    _lib.call("qbn_cls_metrics", _ptr(_f32(probs).contiguous()), _ptr(target.contiguous()), B, K, float(scale), n_bins, _ptr(out), _stream())
    return out


def reg_metrics_accumulate(mean, var, target, out):
    m, v, t = _f32(mean).contiguous().reshape(-1), _f32(var).contiguous().reshape(-1), _f32(target).contiguous().reshape(-1)
    _lib.call("qbn_reg_metrics", _ptr(m), _ptr(v), _ptr(t), m.numel(), _ptr(out), _stream())
    return out


# ------------------------------------------------------------------------------------------------
# glue
# ------------------------------------------------------------------------------------------------
def maxpool2x2(x):
    B, C, H, W = x.shape
    out = torch.empty((B, C, H // 2, W // 2), dtype=torch.float32, device=x.device, memory_format=CL)
    _lib.call("qbn_maxpool2x2", _ptr(x), B, H, W, C, _ptr(out), _stream())
    return out


def avgpool_all(x, divisor=0.0):
    """Mean over the HxW plane; for zero-bordered maps pass divisor = interior size."""
    B, C, H, W = x.shape
    out = torch.empty((B, C), dtype=torch.float32, device=x.device)
    _lib.call("qbn_avgpool_all", _ptr(x), B, H * W, C, float(divisor), _ptr(out), _stream())
    return out


def conv_s1_forward(x, w, n_samples, N, R, S, scale=None, shift=None, residual=None, relu=False, flags=0, w_shared=False, out=None):
    """tcgen05 zero-copy-im2col conv on the zero-bordered layout.  x [n_samples*B, C, Hp, Wp] channels-last
    (Hp = H+R-1), TF32-exact; w [n_samples, N*R*S*C] packed OHWI, TF32-exact.  Returns [n_samples*B, N, Hp, Wp]."""
    SB, C, Hp, Wp = x.shape
    B = SB // n_samples
    if out is None:
        out = torch.empty((SB, N, Hp, Wp), dtype=torch.float32, device=x.device, memory_format=CL)
    _lib.call("qbn_conv_s1_fwd", n_samples, B, Hp, Wp, C, N, R, S, _ptr(x), _ptr(w), int(w_shared), _ptr(scale), _ptr(shift),
              _ptr(residual), int(bool(relu)) | int(flags), _ptr(out), _stream())
    return out


def nchw_to_nhwc(x):
    B, C, H, W = x.shape
    xc = _f32(x).contiguous()
    out = torch.empty((B, C, H, W), dtype=torch.float32, device=x.device, memory_format=CL)
    _lib.call("qbn_nchw_to_nhwc", _ptr(xc), B, C, H * W, _ptr(out), _stream())
    return out


# ------------------------------------------------------------------------------------------------
# planar-C4 path (include/qbn.h "planar-C4 path", csrc/p4_layout.cuh)
# ------------------------------------------------------------------------------------------------
class P4Map:
    """A batch of zero-bordered maps in the planar-C4 layout: buf [C/4][phases * n_img*Hp*Wp + tail][4].
    The zeros are SHARED between neighbours: `bh` zero rows on top of every map (they are the bottom padding of the map
    above), `bw` zero columns on the left of every row (the right padding of the row above), Hp = H + bh, Wp = W + bw, and
    a zero tail of bh*Wp + bw pixels after the last map.  phases == 4: phase-split storage of a 2(Hp-1) x 2(Wp-1) map for
    a stride-2 consumer (phase (a, b) holds the pixels (2i+a, 2j+b))."""
    __slots__ = ("buf", "n_img", "C", "Hp", "Wp", "border", "phases")

    def __init__(self, buf, n_img, C, Hp, Wp, border, phases=1):
        self.buf, self.n_img, self.C, self.Hp, self.Wp, self.border, self.phases = buf, n_img, C, Hp, Wp, border, phases

    @property
    def plane_rows(self):
        return self.buf.stride(0) // 4          # distance between chunk planes (a view of a larger buffer keeps it)

    def images(self, i0, n):
        """View of images [i0, i0 + n) (same planes, pointer offset): an output slice for a launch that covers part of the batch."""
        assert self.phases == 1
        hw = self.Hp * self.Wp
        return P4Map(self.buf[:, i0 * hw:], n, self.C, self.Hp, self.Wp, self.border, 1)

    @staticmethod
    def tail_rows(Hp, Wp, border):
        return border[0] * Wp + border[1]

    @staticmethod
    def empty(n_img, C, Hp, Wp, border, phases=1, device="cuda", zero=True):
        """Always zero-initialised: kernels never write the tail (and never the border of a phase-split map)."""
        shape = (C // 4, phases * n_img * Hp * Wp + P4Map.tail_rows(Hp, Wp, border), 4)
        return P4Map(torch.zeros(shape, dtype=torch.float32, device=device), n_img, C, Hp, Wp, tuple(border), phases)

    @staticmethod
    def from_nchw(x, border, phase_split=False):
        """Layout conversion with torch ops (tests / entry only): x [n_img, C, H, W] unpadded."""
        n, C, H, W = x.shape
        xh = x.permute(0, 2, 3, 1)
        if phase_split:
            parts = [torch.nn.functional.pad(xh[:, a::2, b::2, :], (0, 0, 1, 0, 1, 0)) for a in (0, 1) for b in (0, 1)]
            Hp, Wp = H // 2 + 1, W // 2 + 1
            rows = torch.stack(parts).reshape(4 * n * Hp * Wp, C)
            border, phases = (1, 1), 4
        else:
            bh, bw = border
            Hp, Wp = H + bh, W + bw
            rows = torch.nn.functional.pad(xh, (0, 0, bw, 0, bh, 0)).reshape(n * Hp * Wp, C)
            phases = 1
        rows = torch.nn.functional.pad(rows, (0, 0, 0, P4Map.tail_rows(Hp, Wp, border)))
        buf = rows.reshape(-1, C // 4, 4).permute(1, 0, 2).contiguous()
        return P4Map(buf, n, C, Hp, Wp, tuple(border), phases)

    def to_nchw(self, keep_border=False):
        body = self.phases * self.n_img * self.Hp * self.Wp
        rows = self.buf[:, :body].permute(1, 0, 2).reshape(self.phases, self.n_img, self.Hp, self.Wp, self.C)
        if self.phases == 4:
            H, W = 2 * (self.Hp - 1), 2 * (self.Wp - 1)
            out = torch.empty((self.n_img, H, W, self.C), dtype=self.buf.dtype, device=self.buf.device)
            k = 0
            for a in (0, 1):
                for b in (0, 1):
                    out[:, a::2, b::2, :] = rows[k][:, 1:, 1:, :]
                    k += 1
            return out.permute(0, 3, 1, 2)
        m = rows[0]
        if not keep_border:
            bh, bw = self.border
            m = m[:, bh:, bw:, :]
        return m.permute(0, 3, 1, 2)

    def tail(self):
        return self.buf[:, self.phases * self.n_img * self.Hp * self.Wp:]


def p4_stage_input(x, C_pad, border, phase_split=False):
    """x [n_img, C, H, W] in channels_last memory -> planar-C4 map of tf32(x) with the channels zero-padded to C_pad: borders and
    tail written by the same launch (qbn_p4_stage_input; the entry of the evaluation engines)."""
    n, C, H, W = x.shape
    xh = x.permute(0, 2, 3, 1)
    if not (xh.is_contiguous() and x.dtype == torch.float32):
        xh = xh.float().contiguous()
    if phase_split:
        border = (1, 1)
        Hp, Wp, phases = H // 2 + 1, W // 2 + 1, 4
    else:
        Hp, Wp, phases = H + border[0], W + border[1], 1
    rows = phases * n * Hp * Wp + P4Map.tail_rows(Hp, Wp, border)
    buf = torch.empty((C_pad // 4, rows, 4), dtype=torch.float32, device=x.device)
    _lib.call("qbn_p4_stage_input", _ptr(xh), n, H, W, C, C_pad, border[0], border[1], int(phase_split), rows, _ptr(buf), None, _stream())
    return P4Map(buf, n, C_pad, Hp, Wp, tuple(border), phases)


def p4_weight_floats(C, N, R, S, stride=1):
    n = ctypes.c_longlong(0)
    _lib.call("qbn_p4_weight_floats", C, N, R, S, stride, ctypes.byref(n))
    return int(n.value)


def p4_block_weights(w_ohwi, N, C, taps, stride=1, out=None, cb=0):
    """[n_mats, N*taps*C] packed OHWI -> blocked [n_mats, p4_weight_floats] (the blocking depends on the stride)."""
    w_ohwi = w_ohwi.reshape(-1, N * taps * C).contiguous()
    n_mats = w_ohwi.shape[0]
    R = taps
    if out is None:
        out = torch.empty((n_mats, p4_weight_floats(C, N, R, 1, stride)), dtype=torch.float32, device=w_ohwi.device)
    _lib.call("qbn_p4_block_weights", _ptr(w_ohwi), n_mats, N, C, taps, stride, int(cb), _ptr(out), _stream())
    return out


def sample_weights_blocked(mu_b, sigma_b, N, C, taps, n_samples=1, eps=None, seed=0, layer_id=0, sample0=0, round_tf32=True, out=None, stride=1):
    if out is None:
        out = torch.empty((n_samples, mu_b.numel()), dtype=torch.float32, device=mu_b.device)
    _lib.call("qbn_sample_weights_blocked", _ptr(mu_b), _ptr(sigma_b), N, C, taps, stride, n_samples, _ptr(eps), seed, layer_id, sample0, _ptr(out),
              int(round_tf32), _stream())
    return out


def sample_weights_blocked_multi(jobs_dev, n_jobs, max_floats, n_samples, seed, sample0, round_tf32=True):
    """jobs_dev: uint8 CUDA tensor holding an array of _lib.P4SampleJob (all planar layers of a chunk, one launch)."""
    _lib.call("qbn_sample_weights_blocked_multi", _ptr(jobs_dev), n_jobs, max_floats, n_samples, seed, sample0, int(round_tf32), _stream())


def dropout_masks_multi(jobs_dev, n_jobs, max_elems, n_samples, keep_prob, seed, sample0):
    """jobs_dev: uint8 CUDA tensor holding an array of _lib.MaskJob (all dropout sites of a chunk, one launch)."""
    _lib.call("qbn_dropout_masks_multi", _ptr(jobs_dev), n_jobs, max_elems, n_samples, float(keep_prob), seed, sample0, _stream())


def conv_p4_forward(x, w, n_samples, N, R, S, stride=1, scale=None, shift=None, residual=None, relu=False, flags=0, w_shared=False,
                    out=None, phase_split_out=False, out_mask=None, out_mask_mult=1.0):
    """qbn_conv_p4_fwd.  x: P4Map (phase-split when stride == 2); w: blocked sampled weights [n_samples, ...];
    residual: P4Map with the output geometry.  Returns a P4Map."""
    stacked = bool(int(flags) & QBN_FLAG_X_SHARED_STACKED)
    B = x.n_img if stacked else x.n_img // n_samples
    n_out = x.n_img * n_samples if stacked else x.n_img
    if stride == 2 and x.phases != 4:
        raise _lib.QbnError("stride-2 planar conv needs a phase-split input")
    border = ((R - 1) // 2, (S - 1) // 2) if stride == 1 else (1, 1)
    if out is None:
        if phase_split_out:
            H, W = x.Hp - border[0], x.Wp - border[1]
            out = P4Map.empty(n_out, N, H // 2 + 1, W // 2 + 1, (1, 1), 4, x.buf.device)
        else:
            out = P4Map.empty(n_out, N, x.Hp, x.Wp, border, 1, x.buf.device)
    fl = int(bool(relu)) | int(flags) | (QBN_FLAG_OUT_PHASE_SPLIT if phase_split_out else 0)
    _lib.call("qbn_conv_p4_fwd", n_samples, B, x.Hp, x.Wp, x.C, N, R, S, stride, _ptr(x.buf), x.plane_rows, _ptr(w), int(w_shared), _ptr(scale),
              _ptr(shift), _ptr(residual.buf if residual is not None else None), residual.plane_rows if residual is not None else 0,
              _ptr(out_mask), float(out_mask_mult), fl, _ptr(out.buf), out.plane_rows, _stream())
    return out


def p4_shortcut_block_channels(C, C2):
    return int(_lib.load().qbn_p4_shortcut_block_channels(C, C2))


def conv_p4_shortcut_forward(x, w, x2, n_samples, N, R, S, scale=None, shift=None, relu=False, flags=0, out=None):
    """qbn_conv_p4_shortcut_fwd: stride-1 conv of x plus the 1x1 stride-2 shortcut of the phase-split block input x2 in
    one accumulator.  w: [n_samples, main blocks + shortcut blocks] (both carrying their BatchNorm scale)."""
    B = x.n_img // n_samples
    if x2.phases != 4 or (x2.Hp, x2.Wp, x2.n_img) != (x.Hp, x.Wp, x.n_img):
        raise _lib.QbnError("fused shortcut: x2 must be the phase-split block input with the output geometry")
    if out is None:
        out = P4Map.empty(x.n_img, N, x.Hp, x.Wp, x.border, 1, x.buf.device)
    _lib.call("qbn_conv_p4_shortcut_fwd", n_samples, B, x.Hp, x.Wp, x.C, N, R, S, _ptr(x.buf), x.plane_rows, _ptr(w), _ptr(x2.buf), x2.plane_rows,
              x2.C, _ptr(scale), _ptr(shift), int(bool(relu)) | int(flags), _ptr(out.buf), out.plane_rows, _stream())
    return out


def avgpool_p4(x, divisor):
    out = torch.empty((x.n_img, x.C), dtype=torch.float32, device=x.buf.device)
    _lib.call("qbn_avgpool_p4", _ptr(x.buf), x.n_img, x.Hp * x.Wp, x.plane_rows, x.C, float(divisor), _ptr(out), _stream())
    return out


# ------------------------------------------------------------------------------------------------
# int8 on the planar zero-copy kernel ("planar C16", csrc/p4_layout.cuh)
# ------------------------------------------------------------------------------------------------
def pad32(c):
    return (int(c) + 31) // 32 * 32


class P16Map:
    """A batch of quint8 maps (at most 7 bits) stored as (q - zero_point) int8 in the planar-C16 layout:
    buf [C_pad/16][phases * n_img*Hp*Wp + tail][16], one zero row on top / zero column on the left of every map (border 1),
    C_pad = channels zero-padded to a multiple of 32.  `scale`, `zero_point`: the per-tensor affine parameters of the quint8
    tensor it stands for (torch's q_scale / q_zero_point); phases == 4: phase-split storage for a stride-2 consumer."""
    __slots__ = ("buf", "n_img", "C", "C_pad", "Hp", "Wp", "phases", "scale", "zero_point", "bits")

    def __init__(self, buf, n_img, C, Hp, Wp, phases, scale, zero_point, bits=7):
        self.buf, self.n_img, self.C, self.C_pad, self.Hp, self.Wp, self.phases = buf, n_img, C, buf.shape[0] * 16, Hp, Wp, phases
        self.scale, self.zero_point, self.bits = float(scale), int(zero_point), int(bits)

    @property
    def plane_rows(self):
        return self.buf.stride(0) // 16

    @staticmethod
    def empty(n_img, C, Hp, Wp, phases=1, device="cuda", scale=1.0, zero_point=0, bits=7):
        """Zero-initialised: kernels never write the tail, the padded channel planes, nor the border of a phase-split map."""
        shape = (pad32(C) // 16, phases * n_img * Hp * Wp + Wp + 1, 16)
        return P16Map(torch.zeros(shape, dtype=torch.int8, device=device), n_img, C, Hp, Wp, phases, scale, zero_point, bits)

    @staticmethod
    def from_quint8(q_nhwc, scale, zero_point, bits=7, out=None):
        """q_nhwc: uint8 [n_img, C, H, W] channels-last (QTensor.q) with values in [0, 2^bits - 1], bits <= 7."""
        n, C, H, W = q_nhwc.shape
        if bits > 7:
            raise _lib.QbnError("the planar int8 path holds q - zero_point as int8: activations must be clamped to <= 7 bits")
        q = q_nhwc.contiguous(memory_format=CL)
        m = out if out is not None else P16Map.empty(n, C, H + 1, W + 1, 1, q.device, scale, zero_point, bits)
        m.scale, m.zero_point, m.bits = float(scale), int(zero_point), int(bits)
        _lib.call("qbn_i8_p16_from_nhwc", _ptr(q), n, H, W, C, m.C_pad, int(zero_point), m.plane_rows, _ptr(m.buf), _stream())
        return m

    def to_quint8(self):
        """uint8 [n_img, C, H, W] channels-last (tests / exit of the layout)."""
        if self.phases == 4:
            H2, W2 = self.Hp - 1, self.Wp - 1
            body = 4 * self.n_img * self.Hp * self.Wp
            rows = self.buf[:, :body].permute(1, 0, 2).reshape(4, self.n_img, self.Hp, self.Wp, self.C_pad)
            out = torch.empty((self.n_img, 2 * H2, 2 * W2, self.C_pad), dtype=torch.int16, device=self.buf.device)
            k = 0
            for a in (0, 1):
                for b in (0, 1):
                    out[:, a::2, b::2, :] = rows[k][:, 1:, 1:, :].to(torch.int16)
                    k += 1
            q = (out[..., :self.C] + self.zero_point).to(torch.uint8)
            return q.permute(0, 3, 1, 2)
        H, W = self.Hp - 1, self.Wp - 1
        out = torch.empty((self.n_img, self.C, H, W), dtype=torch.uint8, device=self.buf.device, memory_format=CL)
        _lib.call("qbn_i8_p16_to_nhwc", _ptr(self.buf), self.n_img, H, W, self.C, self.zero_point, self.plane_rows, _ptr(out), _stream())
        return out


def p16_weight_bytes(C_pad, N, R, S, stride):
    out = ctypes.c_longlong(0)
    _lib.call("qbn_p16_weight_bytes", C_pad, N, R, S, stride, ctypes.byref(out))
    return int(out.value)


def i8_p16_block_weights(w, C_pad, stride, out=None):
    """w int8 [n, N, C, R, S] (or [n, N, C]) sampled weights in the sampler's order -> blocked operands [n, bytes]."""
    n, N, C = w.shape[0], w.shape[1], w.shape[2]
    R, S = (w.shape[3], w.shape[4]) if w.dim() == 5 else (1, 1)
    nbytes = p16_weight_bytes(C_pad, N, R, S, stride)
    if out is None:
        out = torch.empty((n, nbytes), dtype=torch.int8, device=w.device)
    _lib.call("qbn_i8_p16_block_weights", _ptr(w.contiguous()), n, N, C, C_pad, R * S, stride, _ptr(out), _stream())
    return out


def i8_conv_p16_forward(x, w_blocked, n_samples, N, R, S, stride, bias, s_w, z_w, s_out, z_out, relu, act_bits, out, residual=None,
                        add_qp=None, add_relu=True, x_shared=False, w_shared=False, out_phase_split=False, acc_dump=None):
    """x, out, residual: P16Map.  Returns `out` with its (scale, zero_point) set to the result's."""
    rq = _lib.I8Requant()
    rq.s_x, rq.s_w, rq.z_w, rq.s_out, rq.z_out = float(x.scale), float(s_w), int(z_w), float(s_out), int(z_out)
    rq.relu, rq.act_max = int(bool(relu)), (1 << act_bits) - 1
    if residual is not None:
        rq.s_res, rq.z_res, rq.s_add, rq.z_add, rq.add_relu = float(residual.scale), int(residual.zero_point), float(add_qp[0]), int(add_qp[1]), int(add_relu)
        if (residual.C * (residual.Hp - 1) * (residual.Wp - 1) * (residual.n_img // max(1, n_samples))) % 64 != 0:
            raise _lib.QbnError("fused quantized::add needs maps whose element count is a multiple of 64 (ATen's vector body)")
    if stride == 2:
        assert x.phases == 4, "a stride-2 planar conv reads a phase-split map"
    Hp, Wp = x.Hp, x.Wp
    B = (x.n_img if x_shared else x.n_img // n_samples)
    _lib.call("qbn_i8_conv_p16_fwd", n_samples, B, Hp, Wp, x.C_pad, N, R, S, stride, _ptr(x.buf), x.plane_rows, int(x_shared), _ptr(w_blocked),
              int(w_shared), _ptr(bias), ctypes.byref(rq), _ptr(residual.buf) if residual is not None else None,
              residual.plane_rows if residual is not None else 0, QBN_FLAG_OUT_PHASE_SPLIT if out_phase_split else 0, _ptr(out.buf),
              out.plane_rows, _ptr(acc_dump), _stream())
    if residual is not None:
        out.scale, out.zero_point = float(add_qp[0]), int(add_qp[1])
    else:
        out.scale, out.zero_point = float(s_out), int(z_out)
    out.bits = act_bits
    return out


def i8_p16_dropout(x, mask, n_samples, B, s_m, z_m, multiplier, act_bits, out, x_shared=False, residual=None, add_qp=None, add_relu=False):
    """int8 MC-Dropout of a chunk on planar-C16 maps (+ the BasicBlock's residual add): include/qbn.h qbn_i8_p16_dropout.
    x, residual, out: P16Map (same geometry and phases); mask: fp32 [n_samples*B, C]; the output's qparams are set here."""
    s_drop = float(s_m) * float(multiplier)              # quantized::mul_scalar: scale * scalar in double
    rq = None
    if residual is not None:
        rq = _lib.I8Requant(1.0, 1.0, 0, 1.0, 0, 0, (1 << act_bits) - 1, float(residual.scale), int(residual.zero_point), float(add_qp[0]), int(add_qp[1]),
                            int(bool(add_relu)))
    if x.phases != 1 or (residual is not None and residual.phases != 1):
        raise _lib.QbnError("qbn_i8_p16_dropout reads maps in the normal layout (the output may be phase-split)")
    _lib.call("qbn_i8_p16_dropout", _ptr(x.buf), x.plane_rows, int(bool(x_shared)), int(n_samples), int(B), x.Hp, x.Wp, int(out.phases == 4), x.C, float(x.scale),
              _ptr(mask), float(s_m), int(z_m), s_drop, (1 << act_bits) - 1, _ptr(residual.buf if residual is not None else None),
              residual.plane_rows if residual is not None else 0, ctypes.byref(rq) if rq is not None else None, _ptr(out.buf), out.plane_rows, _stream())
    if residual is not None:
        out.scale, out.zero_point = float(add_qp[0]), int(add_qp[1])
    else:
        out.scale, out.zero_point = s_drop, int(z_m)
    out.bits = act_bits
    return out


def i8_p16_avgpool(x, act_bits=7):
    """nn.AvgPool2d over the whole map of a P16Map -> uint8 [n_img, C] (same scale / zero point)."""
    out = torch.empty((x.n_img, x.C), dtype=torch.uint8, device=x.buf.device)
    _lib.call("qbn_i8_p16_avgpool", _ptr(x.buf), x.n_img, x.Hp - 1, x.Wp - 1, x.C, x.zero_point, x.plane_rows, 0, (1 << act_bits) - 1,
              _ptr(out), _stream())
    return out
