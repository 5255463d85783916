"""TEST INFRASTRUCTURE: stand-ins for the int8 entry points of `qbn_b200.ops`, backed by the oracle, so that the HOST-side
logic of the int8 path (module plumbing, QTensor flow through the model glue, sample batching, checkpoint loading) can be
exercised on a box without a GPU.  The product never uses these — its ops raise on CPU tensors; only tests install them,
through `emulated_int8_ops(monkeypatch)`.  Semantics follow include/qbn.h entry by entry; layouts follow ops.py."""
import numpy as np
import torch

import oracle.qbn_oracle as O
from oracle import philox


def _nchw_ints(t):
    return t.detach().cpu().numpy().astype(np.int32)            # logical NCHW view, whatever the memory format


def _like_cl(arr, dtype=torch.uint8):
    t = torch.as_tensor(np.ascontiguousarray(arr).astype(np.uint8 if dtype == torch.uint8 else np.int8))
    return t.contiguous(memory_format=torch.channels_last) if t.dim() == 4 else t


def quantize_u8(x, scale, zp, qmin=0, qmax=255):
    return torch.as_tensor(O.quantize(x.detach().float().cpu().numpy(), scale, zp, qmin, qmax).astype(np.uint8))


def dequantize_u8(q, scale, zp):
    return torch.as_tensor(O.dequantize(q.cpu().numpy(), scale, zp))


def i8_sample_weights(mu_q, sigma_q, params, n_samples=1, eps=None, seed=0, layer_id=0, sample0=0):
    n = mu_q.numel()
    mu, sg = mu_q.cpu().numpy().astype(np.int32), sigma_q.cpu().numpy().astype(np.int32)
    bits = {(-128, 127): 8, (-64, 63): 7, (-32, 31): 6, (-16, 15): 5, (-8, 7): 4, (-4, 3): 3, (-2, 1): 2}[(params.w_min, params.w_max)]
    out = np.empty((n_samples, n), np.int8)
    for s in range(n_samples):
        e = eps[s].cpu().numpy().reshape(-1) if eps is not None else philox.philox_normal(n, seed, layer_id, sample0 + s)
        out[s] = O.i8_sample_weight(mu, params.s_mu, params.z_mu, sg, params.s_sigma, params.z_sigma, e, params.s_mul, params.z_mul,
                                    params.s_add, params.z_add, w_bits=bits, n_vec=None if params.n_vec < 0 else params.n_vec)
    return torch.as_tensor(out)


def i8_conv_forward(x_q, s_x, z_x, w_q, s_w, z_w, d, bias, s_out, z_out, relu, act_bits=7, n_samples=1, x_shared=True,
                    w_shared=False, want_acc=False, path=0, linear=False, x_bits=8):
    assert not want_acc
    b = None if bias is None else bias.detach().cpu().numpy()
    outs = []
    for s in range(n_samples):
        w = w_q[0 if w_shared else s].cpu().numpy().astype(np.int32)
        if linear:
            xs = x_q if x_shared else x_q.reshape(n_samples, d.B, -1)[s]
            y, _ = O.i8_linear(xs.cpu().numpy(), s_x, z_x, w.reshape(d.N, d.C), s_w, z_w, b, s_out, z_out, bool(relu), act_bits=act_bits)
        else:
            xs = x_q if x_shared else x_q[s * d.B:(s + 1) * d.B]
            w = w.reshape(d.N, d.R, d.S, d.C).transpose(0, 3, 1, 2)                  # packed OHWI -> OIHW
            assert d.stride_h == d.stride_w and d.pad_h == d.pad_w and d.dil_h == d.dil_w
            y, _ = O.i8_conv(_nchw_ints(xs), s_x, z_x, w, s_w, z_w, b, s_out, z_out, d.stride_h, d.pad_h, d.dil_h, bool(relu), act_bits=act_bits)
        outs.append(y)
    lead = n_samples if (n_samples > 1 or not x_shared) else None
    if linear:
        y = np.stack(outs) if lead is not None else outs[0]
        return torch.as_tensor(y.astype(np.uint8))
    return _like_cl(np.concatenate(outs, 0))


def i8_add(a, sa, za, b, sb, zb, so, zo, act_bits=7, n_vec=-1, relu=False):
    def mem_order(t):                                           # the kernel walks memory: NHWC for channels_last operands
        return t.permute(0, 2, 3, 1).contiguous().numpy().astype(np.int32) if t.dim() == 4 else t.numpy().astype(np.int32)
    y = O.qadd(mem_order(a), sa, za, mem_order(b), sb, zb, so, zo, 0, 255, n_vec=None if n_vec < 0 else n_vec)
    y = np.clip(y, max(0, int(zo)) if relu else 0, O.UINT_BOUNDS[act_bits][1])      # the kernel's output clamp [lo, hi]; relu: lo = zero point
    return _like_cl(y.transpose(0, 3, 1, 2)) if a.dim() == 4 else torch.as_tensor(y.astype(np.uint8))


def i8_relu(x_q, z_x, act_bits=7):
    y = O.i8_relu(_nchw_ints(x_q), z_x, act_bits=act_bits)
    return _like_cl(y) if x_q.dim() == 4 else torch.as_tensor(y.astype(np.uint8))


def i8_avgpool(x_q, z_x, k, act_bits=7):
    return _like_cl(O.i8_avgpool(_nchw_ints(x_q), z_x, k, act_bits=act_bits))


def i8_dropout(x_q, s_x, z_x, p, s_m, z_m, mask=None, key=(0, 0, 0), act_bits=7):
    assert mask is not None, "the CPU stand-in has no Philox: inject the mask"
    m = mask.detach().cpu().numpy()
    x = _nchw_ints(x_q)
    mult = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    y = O.i8_dropout(x, s_x, z_x, m, s_m, z_m, float(mult), act_bits=act_bits)
    y = y[0] if isinstance(y, tuple) else y
    return _like_cl(y) if x_q.dim() == 4 else torch.as_tensor(np.asarray(y).astype(np.uint8))


def mc_mean(probs):
    return probs.float().mean(0)


def emulated_int8_ops(monkeypatch):
    """Install the stand-ins on qbn_b200.ops for the duration of a test (pytest's monkeypatch undoes it)."""
    from qbn_b200 import ops
    for name in ("quantize_u8", "dequantize_u8", "i8_sample_weights", "i8_conv_forward", "i8_add", "i8_relu", "i8_avgpool", "i8_dropout", "mc_mean"):
        monkeypatch.setattr(ops, name, globals()[name])
