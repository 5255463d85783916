"""TEST INFRASTRUCTURE — CPU restatement (oracle) of the reference's stochastic-layer hot path.

Every function restates, in plain torch-CPU fp32 / numpy integer arithmetic, what the cited
reference lines compute (paths relative to the reference repo root).  The float functions call
the same torch CPU operators the reference calls (torch.mm, F.conv2d, F.softplus — the
reference ships no kernels of its own); the integer functions restate the
ATen QuantizedCPU / FBGEMM arithmetic behind torch.ops.quantized.* (third-party, torch pinned
==1.7.1 in requirements.txt:54; executed here with the installed torch 2.11).

PINNING: tests/test_oracle_vs_golden.py checks every function against golden vectors produced
by oracle/make_golden.py, which runs the UNMODIFIED reference modules (oracle/ref_harness.py)
on seeded inputs with replayed noise.  ECE is pinned against the reference's own binning code
(experiments/utils.py:293-304) only up to bin-edge ties because torchmetrics is not installed.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (quantised-bayesian-nets_b200/) never does.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F
from qbn_b200.synthetic import LeNetBBBParams, MLPBBBParams, ResNetBBBParams, resnet_mc_state_dict  # noqa: F401  (seeded parameter containers)

f32 = np.float32
NOISE_SCALE = float(0.02362204724)  # bbb/quantized/__init__.py:1
NOISE_ZERO_POINT = 0  # bbb/quantized/__init__.py:2
UINT_BOUNDS = {8: [0, 255], 7: [0, 127], 6: [0, 63], 5: [0, 31], 4: [0, 15], 3: [0, 7], 2: [0, 3]}  # src/utils.py:18
INT_BOUNDS = {8: [-128, 127], 7: [-64, 63], 6: [-32, 31], 5: [-16, 15], 4: [-8, 7], 3: [-4, 3], 2: [-2, 1]}  # src/utils.py:19-20


def _t(x):
    return x if torch.is_tensor(x) else torch.as_tensor(np.asarray(x))


# ------------------------------------------------------------------------------------------------
# A1/A2  local reparametrisation forward
# ------------------------------------------------------------------------------------------------
def lrt_linear_fwd(x, mu, rho, bias, eps):
    """bbb/linear.py:32-40.  eps has out's shape [B,N] (drawn after both mm's)."""
    x, mu, rho, eps = _t(x), _t(mu), _t(rho), _t(eps)
    mean = torch.mm(x, mu.t())
    std = torch.sqrt(1e-8 + torch.mm(torch.pow(x, 2), torch.pow(F.softplus(rho).t(), 2)))
    b = _t(bias) if bias is not None else 0.0
    return mean + std * eps + b, std


def lrt_conv_fwd(x, mu, rho, bias, eps, stride=1, padding=0, dilation=1):
    """bbb/conv.py:24-32.  x NCHW, mu/rho OIHW, eps NCHW like the output."""
    x, mu, rho, eps = _t(x), _t(mu), _t(rho), _t(eps)
    z_mean = F.conv2d(x, mu, None, stride, padding, dilation, 1)
    z_std = torch.sqrt(1e-8 + F.conv2d(torch.pow(x, 2), torch.pow(F.softplus(rho), 2), None, stride, padding, dilation, 1))
    z = z_mean + z_std * eps
    if bias is not None:
        z = z + _t(bias).reshape(1, -1, 1, 1)
    return z, z_std


# ------------------------------------------------------------------------------------------------
# A3  backward of A1/A2 in closed form (what autograd derives from linear.py:32-40 / conv.py:24-32)
# ------------------------------------------------------------------------------------------------
def lrt_linear_bwd(x, mu, rho, eps, std, g):
    """g = dL/dout [B,N].  Returns dx, dmu, drho, dbias."""
    x, mu, rho, eps, std, g = map(_t, (x, mu, rho, eps, std, g))
    sigma = F.softplus(rho)
    dv = g * eps / (2.0 * std)
    dmu = g.t() @ x
    dsig2 = dv.t() @ (x * x)
    drho = dsig2 * 2.0 * sigma * torch.sigmoid(rho)
    dx = g @ mu + 2.0 * x * (dv @ (sigma * sigma))
    return dx, dmu, drho, g.sum(0)


def lrt_conv_bwd(x, mu, rho, eps, std, g, stride=1, padding=0, dilation=1):
    """Same algebra with conv transposes (torch.nn.grad restates the two adjoints)."""
    x, mu, rho, eps, std, g = map(_t, (x, mu, rho, eps, std, g))
    sigma = F.softplus(rho)
    dv = g * eps / (2.0 * std)
    dmu = torch.nn.grad.conv2d_weight(x, mu.shape, g, stride, padding, dilation, 1)
    dsig2 = torch.nn.grad.conv2d_weight(x * x, mu.shape, dv, stride, padding, dilation, 1)
    drho = dsig2 * 2.0 * sigma * torch.sigmoid(rho)
    dx = torch.nn.grad.conv2d_input(x.shape, mu, g, stride, padding, dilation, 1) + 2.0 * x * torch.nn.grad.conv2d_input(
        x.shape, sigma * sigma, dv, stride, padding, dilation, 1)
    return dx, dmu, drho, g.sum((0, 2, 3))


# ------------------------------------------------------------------------------------------------
# A4  eval-time weight sampling
# ------------------------------------------------------------------------------------------------
def sample_weight(mu, rho, eps):
    """linear.py:43-47 / conv.py:34-38: W = mu + eps*softplus(rho) (mul_noise.mul then add_weight.add)."""
    mu, rho, eps = _t(mu), _t(rho), _t(eps)
    return mu + eps * F.softplus(rho)


def eval_linear_fwd(x, mu, rho, bias, eps):
    """linear.py:42-50."""
    w = sample_weight(mu, rho, eps)
    b = _t(bias) if bias is not None else 0.0
    return torch.mm(_t(x), w.t()) + b


def eval_conv_fwd(x, mu, rho, bias, eps, stride=1, padding=0, dilation=1):
    """conv.py:33-39."""
    w = sample_weight(mu, rho, eps)
    return F.conv2d(_t(x), w, _t(bias) if bias is not None else None, stride, padding, dilation, 1)


# ------------------------------------------------------------------------------------------------
# A5  KL
# ------------------------------------------------------------------------------------------------
def kl_divergence(mu, rho, sigma_prior):
    """utils_bbb.py:3-5 with mu_prior = 0 (linear.py:24-28, conv.py:43-47)."""
    mu, rho = _t(mu), _t(rho)
    sigma = F.softplus(rho)
    sp = torch.ones_like(rho) * float(sigma_prior)
    return 0.5 * (2 * torch.log(sp / sigma) - 1 + (sigma / sp).pow(2) + ((0.0 - mu) / sp).pow(2)).sum()


def kl_grads(mu, rho, sigma_prior):
    """d KL / d mu = mu/sp^2 ; d KL / d rho = (sigma/sp^2 - 1/sigma) * sigmoid(rho)."""
    mu, rho = _t(mu).double(), _t(rho).double()
    sigma = F.softplus(rho)
    sp2 = float(sigma_prior) ** 2
    return (mu / sp2).float(), ((sigma / sp2 - 1.0 / sigma) * torch.sigmoid(rho)).float()


# ------------------------------------------------------------------------------------------------
# A7  fake quantisation with MovingAverageMinMaxObserver
# ------------------------------------------------------------------------------------------------
def observer_update(state, x, averaging_const=0.01):
    """torch/ao/quantization/observer.py:668-683.  state = [min, max, initialised]."""
    x = _t(x)
    mn, mx = float(x.min()), float(x.max())
    if not state[2]:
        return [f32(mn), f32(mx), True]
    c = f32(averaging_const)
    return [f32(state[0] + c * (f32(mn) - state[0])), f32(state[1] + c * (f32(mx) - state[1])), True]


def calc_qparams(mn, mx, qmin, qmax):
    """observer.py:374-410, per_tensor_affine: scale=(max+ - min-)/(qmax-qmin) >= eps_fp32,
    zp = clamp(qmin - round(min-/scale), qmin, qmax)."""
    mn_neg = min(f32(mn), f32(0.0))
    mx_pos = max(f32(mx), f32(0.0))
    scale = f32(f32(mx_pos - mn_neg) / f32(float(qmax - qmin)))
    scale = max(scale, f32(np.finfo(np.float32).eps))
    zp = qmin - int(np.rint(f32(mn_neg / scale)))
    zp = int(min(max(zp, qmin), qmax))
    return f32(scale), zp


def fake_quant(x, scale, zp, qmin, qmax):
    """fake_quantize_per_tensor_affine: y=(clamp(rint(x*(1/s))+z,qmin,qmax)-z)*s, mask = unclamped."""
    x = np.asarray(x, dtype=f32)
    inv = f32(1.0) / f32(scale)
    q = np.rint(x * inv) + zp
    mask = (q >= qmin) & (q <= qmax)
    y = ((np.clip(q, qmin, qmax) - zp).astype(f32) * f32(scale)).astype(f32)
    return y, mask


# ------------------------------------------------------------------------------------------------
# A6  true int8 path — ATen QuantizedCPU / FBGEMM arithmetic restated in numpy
# ------------------------------------------------------------------------------------------------
def quantize(x, scale, zp, qmin, qmax):
    """torch.quantize_per_tensor: q = clamp(rint(x * fp32(1/scale)) + zp, qmin, qmax)
    (conv_q.py:115, linear_q.py:88, dropout.py:34)."""
    x = np.asarray(x, dtype=f32)
    inv = f32(1.0) / f32(scale)
    return np.clip(np.rint(x * inv) + zp, qmin, qmax).astype(np.int32)


def dequantize(q, scale, zp):
    return ((np.asarray(q, dtype=np.int32) - zp).astype(f32) * f32(scale)).astype(f32)


def qmul_multiplier(sa, sb, so):
    """ATen qmul: float multiplier = self_scale * other_scale * (1.0f / out_scale), all fp32."""
    return f32(f32(f32(sa) * f32(sb)) * f32(f32(1.0) / f32(so)))


def qmul(a, sa, za, b, sb, zb, so, zo, qmin=-128, qmax=127):
    """torch.ops.quantized.mul (linear_q.py:91 `self.mul_noise.mul(std, noise)`):
    c=(a-za)(b-zb) int32; r=clamp(rint(fp32(c)*multiplier)+zo)."""
    c = (np.asarray(a, np.int32) - za) * (np.asarray(b, np.int32) - zb)
    m = qmul_multiplier(sa, sb, so)
    return np.clip(np.rint(c.astype(f32) * m) + zo, qmin, qmax).astype(np.int32)


def _fma32(a, b, c):
    # fp32 fused multiply-add (products of two fp32 are exact in fp64)
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f32)


def qadd(a, sa, za, b, sb, zb, so, zo, qmin=-128, qmax=127, relu=False, n_vec=None):
    """torch.ops.quantized.add (linear_q.py:91 `self.add_weight.add`, src/utils.py:55):
    vector body (ATen Vectorized<qint8>::dequantize): d = fma(scale, float(q), -zp*scale);
    scalar tail (< one 64-lane vector at the end): d = (q - zp) * scale.
    f = da + db ; q = clamp(rint(f * fp32(1/so)) + zo)."""
    a = np.asarray(a, np.int32)
    b = np.asarray(b, np.int32)
    shape = a.shape
    a = a.reshape(-1)
    b = b.reshape(-1)
    n = a.size
    if n_vec is None:
        n_vec = (n // 64) * 64
    sa32, sb32 = f32(sa), f32(sb)
    da = _fma32(sa32, a.astype(f32), f32(sa32 * f32(-za)))
    db = _fma32(sb32, b.astype(f32), f32(sb32 * f32(-zb)))
    da[n_vec:] = (a[n_vec:] - za).astype(f32) * sa32
    db[n_vec:] = (b[n_vec:] - zb).astype(f32) * sb32
    fsum = (da + db).astype(f32)
    if relu:
        fsum = np.maximum(fsum, f32(0.0))
    inv = f32(1.0) / f32(so)
    return np.clip(np.rint(fsum * inv) + zo, qmin, qmax).astype(np.int32).reshape(shape)


def i8_sample_weight(mu_q, s_mu, z_mu, sigma_q, s_sigma, z_sigma, eps, s_mul, z_mul, s_add, z_add, w_bits=8, n_vec=None):
    """SURVEY §8a row A6 steps 1-4 = linear_q.py:86-92 / conv_q.py:113-119:
    noise -> quantize_per_tensor(NOISE_SCALE, 0, qint8) -> mul_noise.mul(std, noise) ->
    add_weight.add(weight, .) -> clamp_weight (src/utils.py:32-37)."""
    eps_q = quantize(eps, NOISE_SCALE, NOISE_ZERO_POINT, -128, 127)
    r = qmul(sigma_q, s_sigma, z_sigma, eps_q, NOISE_SCALE, NOISE_ZERO_POINT, s_mul, z_mul)
    w = qadd(mu_q, s_mu, z_mu, r, s_mul, z_mul, s_add, z_add, n_vec=n_vec)
    lo, hi = INT_BOUNDS[w_bits]
    return np.clip(w, lo, hi).astype(np.int32)


def requant_params(s_x, s_w, s_out):
    """qlinear/qconv (ATen, fbgemm): act_times_w = fp32(s_x)*fp32(s_w); mult = act_times_w / fp32(s_out)."""
    atw = f32(f32(s_x) * f32(s_w))
    return atw, f32(atw / f32(s_out))


def requantize(acc, bias, s_x, s_w, s_out, z_out, relu, act_bits=8):
    """FBGEMM ReQuantizeOutput with float bias (A6 step 6) followed by clamp_activation
    (src/utils.py:25-30): y = clamp(rint((fp32(acc) + bias/(sx*sw)) * (sx*sw/so)) + zo, lo, hi)."""
    atw, mult = requant_params(s_x, s_w, s_out)
    xf = np.asarray(acc, np.int32).astype(f32)
    if bias is not None:
        xf = (xf + (np.asarray(bias, f32) / atw).astype(f32)).astype(f32)
    q = np.rint((xf * mult).astype(f32)) + z_out
    lo = z_out if relu else 0
    q = np.clip(q, lo, 255)
    amin, amax = UINT_BOUNDS[act_bits]
    return np.clip(q, amin, amax).astype(np.int32)


def i8_linear(x_q, s_x, z_x, w_q, s_w, z_w, bias, s_out, z_out, relu=False, act_bits=8):
    """torch.ops.quantized.linear[_relu] (linear_q.py:93-94,168-172). x_q [B,K] u8, w_q [N,K] s8."""
    acc = (np.asarray(x_q, np.int64) - z_x) @ (np.asarray(w_q, np.int64) - z_w).T
    acc = acc.astype(np.int32)
    return requantize(acc, bias, s_x, s_w, s_out, z_out, relu, act_bits), acc


def i8_conv(x_q, s_x, z_x, w_q, s_w, z_w, bias, s_out, z_out, stride=1, padding=0, dilation=1, relu=False, act_bits=8):
    """torch.ops.quantized.conv2d[_relu] (conv_q.py:120-125,206-209).  x_q NCHW u8, w_q OIHW s8.
    Zero padding pads with the zero point, i.e. (x - z_x) = 0 outside the image."""
    xi = torch.as_tensor(np.asarray(x_q, np.float64) - z_x)
    wi = torch.as_tensor(np.asarray(w_q, np.float64) - z_w)
    acc = F.conv2d(xi, wi, None, stride, padding, dilation, 1).numpy()  # exact in fp64 (|acc| < 2^31)
    acc = np.rint(acc).astype(np.int32)
    b = None if bias is None else np.asarray(bias, f32).reshape(1, -1, 1, 1)
    return requantize(acc, b, s_x, s_w, s_out, z_out, relu, act_bits), acc


def i8_add(a, sa, za, b, sb, zb, so, zo, act_bits=8, n_vec=None):
    """QFunctional.add on quint8 (src/utils.py:49-55) + clamp_activation."""
    q = qadd(a, sa, za, b, sb, zb, so, zo, 0, 255, n_vec=n_vec)
    amin, amax = UINT_BOUNDS[act_bits]
    return np.clip(q, amin, amax).astype(np.int32)


def i8_relu(x_q, z_x, act_bits=8):
    """torch.relu on a quint8 tensor (BasicBlock.end, models_bbb.py:186) + clamp_activation: ints below the zero
    point are raised to it; qparams unchanged."""
    amin, amax = UINT_BOUNDS[act_bits]
    return np.clip(np.maximum(np.asarray(x_q, np.int32), z_x), amin, amax).astype(np.int32)


def i8_avgpool(x_q, z_x, k, act_bits=8):
    """nn.AvgPool2d(k) on a quint8 NCHW tensor (models_bbb.py:211; ATen qavg_pool2d, output qparams = input's):
    acc = sum(window) - k*k*z int32; q = clamp(rint(fp32(acc) * fp32(1/(k*k))) + z, 0, 255), then clamp_activation."""
    x = np.asarray(x_q, np.int32)
    B, C, H, W = x.shape
    win = x.reshape(B, C, H // k, k, W // k, k).sum(axis=(3, 5)) - k * k * z_x
    q = np.clip(np.rint(win.astype(f32) * f32(f32(1.0) / f32(k * k))) + z_x, 0, 255)
    amin, amax = UINT_BOUNDS[act_bits]
    return np.clip(q, amin, amax).astype(np.int32)


def i8_dropout(x_q, s_x, z_x, mask, s_m, z_m, multiplier, act_bits=8):
    """dropout.py:31-39 in the int8 model: mask -> quint8 at (s_m,z_m); quantized.mul with the
    output at (s_m,z_m); mul_scalar keeps the ints and multiplies the scale by `multiplier`.
    x_q NCHW (or [B,C]); mask [B,C].  Returns (ints, new_scale, z_m)."""
    x_q = np.asarray(x_q, np.int32)
    m_q = quantize(mask, s_m, z_m, 0, 255)
    if x_q.ndim > 2:
        m_q = m_q.reshape(m_q.shape[0], m_q.shape[1], 1, 1)
    c = (x_q - z_x) * (m_q - z_m)
    mult = qmul_multiplier(s_x, s_m, s_m)
    q = np.clip(np.rint(c.astype(f32) * mult) + z_m, 0, 255).astype(np.int32)
    amin, amax = UINT_BOUNDS[act_bits]
    return np.clip(q, amin, amax), float(s_m) * float(multiplier), z_m


# ------------------------------------------------------------------------------------------------
# A8  MC-Dropout (float)
# ------------------------------------------------------------------------------------------------
def dropout_fwd(x, mask, p):
    """dropout.py:15-40: y = x*mask*1/(1-p); mask [B,C] broadcast over H,W for 4-D inputs."""
    x, mask = _t(x), _t(mask)
    if x.dim() > 2:
        mask = mask.view(mask.shape[0], mask.shape[1], 1, 1)
    mult = (torch.ones(1) / (1.0 - torch.ones(1) * p))
    return x * mask * mult


# ------------------------------------------------------------------------------------------------
# A9  Monte-Carlo aggregation
# ------------------------------------------------------------------------------------------------
def mc_mean_probs(probs_list):
    """experiments/utils.py:355: torch.stack(y, dim=1).mean(dim=1)."""
    return torch.stack([_t(p) for p in probs_list], dim=1).mean(dim=1)


def reg_mc_reduce(mu_list, var_list):
    """experiments/utils.py:349-353."""
    mu = torch.stack([_t(m) for m in mu_list], dim=1)
    var = torch.stack([_t(v) for v in var_list], dim=1)
    return mu.mean(dim=1), mu.var(dim=1) + var.mean(dim=1)


# ------------------------------------------------------------------------------------------------
# A10  metrics
# ------------------------------------------------------------------------------------------------
def cls_metric_sums(probs, target, n_bins=10):
    """src/metrics.py:20-29 (error), 48-57 (nll), 76-85 (brier), 104-112 (entropy) as sums, and
    the ECE bin statistics (metrics.py:381-383: 10 equal-width bins on the max-prob confidence,
    norm l1).  Returns dict of python floats + bins array [n_bins,3] = (conf_sum, acc_sum, count)."""
    p = _t(probs).float()
    t = _t(target).long()
    pred = torch.argmax(p, dim=1)
    onehot = F.one_hot(t, num_classes=p.shape[1]).float()
    out = {
        "error": float(torch.sum(pred != t)),
        "nll": float(torch.sum(-onehot * torch.log(p + 1e-8))),
        "brier": float(torch.sum((p - onehot) ** 2)),
        "entropy": float(torch.sum(-p * torch.log(p + 1e-8))),
    }
    conf = p.max(dim=1).values.numpy()
    acc = (pred == t).numpy().astype(np.float64)
    # torchmetrics bucketize(conf, linspace(0,1,n+1), right=True)-1 : bin b = (b/n, (b+1)/n]... with
    # right=True the boundaries satisfy bounds[i-1] <= v < bounds[i]; i.e. bins are [lo, hi).
    bounds = np.linspace(0, 1, n_bins + 1, dtype=np.float32)
    idx = np.clip(np.searchsorted(bounds, conf, side="right") - 1, 0, n_bins - 1)
    bins = np.zeros((n_bins, 3), dtype=np.float64)
    for b in range(n_bins):
        m = idx == b
        bins[b] = (conf[m].astype(np.float64).sum(), acc[m].sum(), m.sum())
    out["bins"] = bins
    return out


def ece_from_bins(bins):
    total = bins[:, 2].sum()
    ece = 0.0
    for conf_sum, acc_sum, cnt in bins:
        if cnt > 0:
            ece += abs(acc_sum / cnt - conf_sum / cnt) * cnt / total
    return ece


def reg_metric_sums(mean, var, target):
    """src/metrics.py:135-157 (gaussian nll), 176-187 (mse), 214-225 (mae) as sums."""
    m, v, t = _t(mean).float().squeeze(), _t(var).float().squeeze(), _t(target).float().squeeze()
    nll = torch.sum(0.5 * torch.log(2 * math.pi * v + 1e-8) + (t - m) ** 2 / (2 * v + 1e-8))
    return {"nll": float(nll), "se": float(((m - t) ** 2).sum()), "ae": float((m - t).abs().sum())}


# ------------------------------------------------------------------------------------------------
# whole-network CPU port used as bench.py's cpu_baseline / --impl reference leg
# ------------------------------------------------------------------------------------------------
def resnet_bbb_eval_forward(P, x, eps_fn):
    """One eval-mode forward of models_bbb.py:226-245 (ConvNetwork_ResNet.forward) with
    conv.py:33-39 / linear.py:42-50 weight sampling.  eps_fn(name, shape) supplies the noise
    (torch.randn for timing; injected tensors for parity)."""

    def conv(name, h, stride, pad):
        mu, rho = P.convs[name]
        return eval_conv_fwd(h, mu, rho, None, eps_fn(name, mu.shape), stride, pad, 1)

    def bn(name, h):
        w, b, rm, rv, eps = P.bns[name]
        return F.batch_norm(h, rm, rv, w, b, False, 0.0, eps)

    h = F.relu(bn("layers.1", conv("layers.0", x, 1, 1)))
    for p, stride, has_sc in P.blocks:
        out = F.relu(bn(p + ".stem.1", conv(p + ".stem.0", h, stride, 1)))
        out = bn(p + ".stem.4", conv(p + ".stem.3", out, 1, 1))
        sc = bn(p + ".shortcut.1", conv(p + ".shortcut.0", h, stride, 0)) if has_sc else h
        h = F.relu(out + sc)
    h = F.avg_pool2d(h, 4).reshape(h.size(0), -1)
    mu, rho = P.fc
    logits = eval_linear_fwd(h, mu, rho, None, eps_fn("fc", mu.shape))
    return F.softmax(logits, dim=-1)


def resnet_bbb_mc_predict(P, x, n_samples, eps_fn=None):
    """experiments/utils.py:342-355: S sequential forwards, stack, mean(dim=1)."""
    if eps_fn is None:
        eps_fn = lambda name, shape: torch.empty(shape).normal_()  # noqa: E731
    ys = [resnet_bbb_eval_forward(P, x, eps_fn) for _ in range(n_samples)]
    return mc_mean_probs(ys)


def resnet_noise_plan(P):
    """Forward order of the BBB layers of ConvNetwork_ResNet (models_bbb.py:226-245; inside a
    BasicBlock the stem runs before the shortcut, :170-178).  -> [(name, weight_shape)]"""
    plan = [("layers.0", tuple(P.convs["layers.0"][0].shape))]
    for p, _, has_sc in P.blocks:
        plan.append((p + ".stem.0", tuple(P.convs[p + ".stem.0"][0].shape)))
        plan.append((p + ".stem.3", tuple(P.convs[p + ".stem.3"][0].shape)))
        if has_sc:
            plan.append((p + ".shortcut.0", tuple(P.convs[p + ".shortcut.0"][0].shape)))
    plan.append(("fc", tuple(P.fc[0].shape)))
    return plan


def replay_noise(seed, shapes):
    """The reference draws `tensor.new(shape).normal_()` from the global CPU generator in forward
    order (linear.py:36-37,44-45; conv.py:28-29,34-35): re-seeding and drawing the same shapes in
    the same order reproduces its noise exactly."""
    torch.manual_seed(seed)
    return [torch.empty(tuple(s)).normal_() for s in shapes]


def lenet_bbb_eval_forward(P, x, eps_fn):
    """ConvNetwork_LeNet.forward (models_bbb.py:120-133) in eval mode."""
    mu, rho = P.layers["layers.0"]
    h = F.max_pool2d(eval_conv_fwd(x, mu, rho, None, eps_fn("layers.0", mu.shape), 1, 2, 1), 2, 2)
    mu, rho = P.layers["layers.2"]
    h = F.max_pool2d(eval_conv_fwd(h, mu, rho, None, eps_fn("layers.2", mu.shape), 1, 2, 1), 2, 2)
    h = h.reshape(h.size(0), -1)
    mu, rho = P.layers["layers.5"]
    h = F.relu(eval_linear_fwd(h, mu, rho, None, eps_fn("layers.5", mu.shape)))
    mu, rho = P.layers["layers.7"]
    return F.softmax(eval_linear_fwd(h, mu, rho, None, eps_fn("layers.7", mu.shape)), dim=-1)


def mlp_bbb_eval_forward(P, x, eps_fn):
    """LinearNetwork.forward (models_bbb.py:61-78): returns (mu, exp(log_var))."""
    h = x
    for name in ("layers.0", "layers.2", "layers.4"):
        mu, rho, b = P.layers[name]
        h = F.relu(eval_linear_fwd(h, mu, rho, b, eps_fn(name, mu.shape)))
    mu, rho, b = P.layers["mu"]
    out_mu = eval_linear_fwd(h, mu, rho, b, eps_fn("mu", mu.shape))
    mu, rho, b = P.layers["log_var"]
    out_lv = eval_linear_fwd(h, mu, rho, b, eps_fn("log_var", mu.shape))
    return out_mu, out_lv.exp()


# ------------------------------------------------------------------------------------------------
# MC-Dropout networks (models_mc.py): deterministic weights, Bernoulli masks per (image, channel)
# ------------------------------------------------------------------------------------------------
def resnet_mc_forward(P, x, mask_fn, p):
    """One forward of mcdropout/models_mc.py:159-226 (ConvNetwork_ResNet) with the BasicBlock of :117-157.
    P: ResNetBBBParams (only the mu tensors are used as the nn.Conv2d / nn.Linear weights).
    mask_fn(name, shape) supplies the Bernoulli(1-p) masks in the reference's draw order (dropout.py:19-30):
    after layers.0-2, then per block stem.3 (after conv-bn-relu), stem.6 (after conv-bn), shortcut.2."""

    def conv(name, h, stride, pad):
        return F.conv2d(h, P.convs[name][0], None, stride, pad)

    def bn(name, h):
        w, b, rm, rv, eps = P.bns[name]
        return F.batch_norm(h, rm, rv, w, b, False, 0.0, eps)

    def drop(name, h):
        return dropout_fwd(h, mask_fn(name, tuple(h.shape[:2])), p)

    h = drop("layers.3", F.relu(bn("layers.1", conv("layers.0", x, 1, 1))))
    for pfx, stride, has_sc in P.blocks:
        out = drop(pfx + ".stem.3", F.relu(bn(pfx + ".stem.1", conv(pfx + ".stem.0", h, stride, 1))))
        out = drop(pfx + ".stem.6", bn(pfx + ".stem.4", conv(pfx + ".stem.3", out, 1, 1)))
        sc = drop(pfx + ".shortcut.2", bn(pfx + ".shortcut.1", conv(pfx + ".shortcut.0", h, stride, 0))) if has_sc else h
        h = F.relu(out + sc)
    h = F.avg_pool2d(h, 4).reshape(h.size(0), -1)
    return F.softmax(F.linear(h, P.fc[0]), dim=-1)


def resnet_mc_mask_plan(P, batch):
    """Draw order and shapes of the dropout masks of one forward -> [(name, (B, C))]."""
    plan = [("layers.3", (batch, 24))]
    for pfx, _, has_sc in P.blocks:
        c = P.convs[pfx + ".stem.0"][0].shape[0]
        plan += [(pfx + ".stem.3", (batch, c)), (pfx + ".stem.6", (batch, c))]
        if has_sc:
            plan.append((pfx + ".shortcut.2", (batch, c)))
    return plan


def replay_masks(seed, shapes, p):
    """dropout.py:21-30: `torch.zeros(shape).bernoulli_(1 - self.p)` with self.p a 1-element tensor, global generator."""
    torch.manual_seed(seed)
    keep = 1.0 - torch.ones(1) * p
    return [torch.empty(tuple(s)).bernoulli_(keep) for s in shapes]


def lenet_mc_forward(P, x, mask_fn, p):
    """mcdropout/models_mc.py:75-110 (ConvNetwork_LeNet): conv5x5-drop-pool, conv5x5-drop-pool, fc-relu-drop-fc.
    P: LeNetBBBParams (mu tensors as the deterministic weights)."""
    w0, w1, w2, w3 = (P.layers[k][0] for k in ("layers.0", "layers.2", "layers.5", "layers.7"))   # BBB container keys
    h = F.conv2d(x, w0, None, 1, 2)
    h = F.max_pool2d(dropout_fwd(h, mask_fn("layers.1", tuple(h.shape[:2])), p), 2, 2)
    h = F.conv2d(h, w1, None, 1, 2)
    h = F.max_pool2d(dropout_fwd(h, mask_fn("layers.4", tuple(h.shape[:2])), p), 2, 2)
    h = F.relu(F.linear(h.reshape(h.size(0), -1), w2))
    h = dropout_fwd(h, mask_fn("layers.9", tuple(h.shape)), p)
    return F.softmax(F.linear(h, w3), dim=-1)
