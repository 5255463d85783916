"""GPU (pytest -m gpu): the tcgen05/TMEM kernels against the fp32 CUDA-core kernels and the oracle.
TF32 tolerance per north_star: rtol 1e-3 (atol 1e-3 * max|ref|: TF32 keeps 10 mantissa bits of each
operand, accumulation is fp32).  int8: bit exact against the IMAD kernel and the oracle."""
import numpy as np
import pytest
import torch

import oracle.qbn_oracle as O

pytestmark = pytest.mark.gpu

SHAPES = [
    # B, C, H, N, k, stride, pad
    (2, 24, 32, 24, 3, 1, 1),     # ResNet layer1
    (2, 24, 32, 48, 3, 2, 1),     # layer2 stride-2 entry
    (2, 24, 32, 48, 1, 2, 0),     # 1x1 stride-2 shortcut
    (2, 48, 16, 48, 3, 1, 1),
    (3, 96, 8, 96, 3, 1, 1),
    (3, 96, 8, 192, 3, 2, 1),
    (5, 192, 4, 192, 3, 1, 1),    # M = 80 < 128 (ragged tile)
    (2, 20, 14, 50, 5, 1, 2),     # LeNet conv2
    (3, 8, 9, 10, 3, 1, 1),       # N=10 -> padded to 16, odd sizes
    (2, 4, 7, 7, 3, 1, 0),        # K = 36 (partial K block), N=7
]


def close(got, ref, rtol, atol_rel):
    got, ref = got.detach().float().cpu().numpy(), ref.detach().float().cpu().numpy()
    atol = atol_rel * max(1e-30, float(np.abs(ref).max()))
    np.testing.assert_allclose(got, ref, rtol=rtol, atol=atol)


@pytest.fixture(scope="module", autouse=True)
def _lib():
    import __graft_entry__ as g
    g.build()


@pytest.mark.parametrize("shape", SHAPES)
def test_tf32_eval_conv(shape):
    from qbn_b200 import ops
    B, C, H, N, k, stride, pad = shape
    g = torch.Generator().manual_seed(hash(shape) & 0xFFFF)
    S = 3
    x = torch.randn(S * B, C, H, H, generator=g).cuda().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(S, N, k, k, C, generator=g) / (C * k * k) ** 0.5).cuda()
    scale = (torch.rand(N, generator=g) + 0.5).cuda()
    shift = torch.randn(N, generator=g).cuda()
    d = ops.make_desc(B, H, H, C, N, k, k, stride, pad, 1)
    res = torch.randn(S * B, N, d.Ho, d.Wo, generator=g).cuda().contiguous(memory_format=torch.channels_last)
    wf = w.reshape(S, -1).contiguous()
    ref = ops.conv_forward(x, wf, d, S, False, False, scale, shift, res, True, None, 1.0, ops.QBN_MATH_FP32)
    got = ops.conv_forward(x, wf, d, S, False, False, scale, shift, res, True, None, 1.0, ops.QBN_MATH_TF32)
    close(got, ref, 1e-3, 1e-3)
    # independent check of sample 1 against torch's own conv (oracle arithmetic)
    xs = x[B:2 * B].cpu()
    yo = torch.nn.functional.conv2d(xs, w[1].permute(0, 3, 1, 2).cpu(), None, stride, pad)
    yo = torch.relu(yo * scale.cpu().view(1, -1, 1, 1) + shift.cpu().view(1, -1, 1, 1) + res[B:2 * B].cpu())
    close(got[B:2 * B], yo, 1e-3, 1e-3)
    # cp.async operand path: input already TF32-exact (what a previous TF32 layer's epilogue writes)
    xt = x.clone()
    xt_i = xt.view(torch.int32)
    xt_i.add_(0x1000).bitwise_and_(~0x1FFF)                       # RNA to 10 mantissa bits (ties away, sign-magnitude)
    ref3 = ops.conv_forward(xt, wf, d, S, False, False, scale, shift, res, True, None, 1.0, ops.QBN_MATH_FP32)
    got3 = ops.conv_forward(xt, wf, d, S, False, False, scale, shift, res, True, None, 1.0, ops.QBN_MATH_TF32, None,
                            ops.QBN_FLAG_A_TF32_READY | ops.QBN_FLAG_OUT_ROUND_TF32)
    close(got3, ref3, 1e-3, 1e-3)
    assert int((got3.view(torch.int32) & 0x1FFF).abs().max()) == 0     # stored activations are TF32-exact
    # shared input / shared weights variants
    got2 = ops.conv_forward(x[:B].contiguous(memory_format=torch.channels_last), wf, d, S, True, False, None, None, None, False, None, 1.0,
                            ops.QBN_MATH_TF32)
    ref2 = ops.conv_forward(x[:B].contiguous(memory_format=torch.channels_last), wf, d, S, True, False, None, None, None, False, None, 1.0,
                            ops.QBN_MATH_FP32)
    close(got2, ref2, 1e-3, 1e-3)


@pytest.mark.parametrize("shape", SHAPES)
def test_tf32_lrt_forward(shape):
    from qbn_b200 import ops
    B, C, H, N, k, stride, pad = shape
    g = torch.Generator().manual_seed(1 + (hash(shape) & 0xFFFF))
    x = torch.randn(B, C, H, H, generator=g).cuda().contiguous(memory_format=torch.channels_last)
    mu = (torch.randn(N, C, k, k, generator=g) / (C * k * k) ** 0.5).cuda()
    rho = torch.empty(N, C, k, k).uniform_(-5, -2, generator=g).cuda()
    bias = torch.randn(N, generator=g).cuda()
    d = ops.make_desc(B, H, H, C, N, k, k, stride, pad, 1)
    eps = torch.randn(B, N, d.Ho, d.Wo, generator=g).cuda().contiguous(memory_format=torch.channels_last)
    p = ops.weight_prep(mu, rho, want=("mu", "sigma2"))
    ref, std_ref = ops.lrt_forward(x, p["mu"], p["sigma2"], bias, d, eps, (0, 0, 0), ops.QBN_MATH_FP32)
    got, std = ops.lrt_forward(x, p["mu"], p["sigma2"], bias, d, eps, (0, 0, 0), ops.QBN_MATH_TF32)
    close(std, std_ref, 1e-3, 1e-3)
    close(got, ref, 1e-3, 1e-3)
    yo, so = O.lrt_conv_fwd(x.cpu(), mu.cpu(), rho.cpu(), bias.cpu(), eps.cpu(), stride, pad)
    close(got, yo, 1e-3, 1e-3)
    # Philox epilogue draws the same stream in both math modes
    r1, _ = ops.lrt_forward(x, p["mu"], p["sigma2"], None, d, None, (5, 6, 7), ops.QBN_MATH_FP32)
    r2, _ = ops.lrt_forward(x, p["mu"], p["sigma2"], None, d, None, (5, 6, 7), ops.QBN_MATH_TF32)
    close(r2, r1, 1e-3, 2e-3)


@pytest.mark.parametrize("shape", [s for s in SHAPES if s[1] % 8 == 0])
def test_i8_umma_bit_exact(shape):
    from qbn_b200 import ops
    B, C, H, N, k, stride, pad = shape
    rng = np.random.default_rng(hash(shape) & 0xFFFF)
    S = 2
    x = torch.as_tensor(rng.integers(0, 128, (S * B, C, H, H)).astype(np.uint8)).cuda().contiguous(memory_format=torch.channels_last)
    w = torch.as_tensor(rng.integers(-128, 128, (S, N, k, k, C)).astype(np.int8)).cuda().reshape(S, -1).contiguous()
    bias = torch.as_tensor(rng.normal(0, 0.5, N).astype(np.float32)).cuda()
    d = ops.make_desc(B, H, H, C, N, k, k, stride, pad, 1)
    s_x, z_x, s_w, z_w, s_o, z_o = 0.021, 17, 0.0037, -3, 0.09, 40
    for relu in (0, 1):
        ref, acc_ref = ops.i8_conv_forward(x, s_x, z_x, w, s_w, z_w, d, bias, s_o, z_o, relu, 7, S, False, False, True, path=1)
        got, acc = ops.i8_conv_forward(x, s_x, z_x, w, s_w, z_w, d, bias, s_o, z_o, relu, 7, S, False, False, True, path=2)
        assert torch.equal(acc, acc_ref)
        assert torch.equal(got, ref)
    # oracle check of sample 0
    wq = w[0].reshape(N, k, k, C).permute(0, 3, 1, 2).cpu().numpy()
    yo, acco = O.i8_conv(x[:B].cpu().numpy(), s_x, z_x, wq, s_w, z_w, bias.cpu().numpy(), s_o, z_o, stride, pad, 1, True, act_bits=7)
    assert np.array_equal(got[:B].cpu().numpy(), yo)
    assert np.array_equal(acc[:B].cpu().numpy(), acco)


def test_tf32_dropout_mask_in_operand_load():
    from qbn_b200 import ops
    B, C, H, N, S = 4, 24, 8, 24, 2
    g = torch.Generator().manual_seed(3)
    x = torch.randn(S * B, C, H, H, generator=g).cuda().contiguous(memory_format=torch.channels_last)
    w = torch.randn(S, N * 9 * C, generator=g).cuda() * 0.05
    mask = (torch.rand(S * B, C, generator=g) < 0.85).float().cuda()
    d = ops.make_desc(B, H, H, C, N, 3, 3, 1, 1, 1)
    ref = ops.conv_forward(x, w, d, S, False, False, None, None, None, False, mask, 1.0 / 0.85, ops.QBN_MATH_FP32)
    got = ops.conv_forward(x, w, d, S, False, False, None, None, None, False, mask, 1.0 / 0.85, ops.QBN_MATH_TF32)
    close(got, ref, 1e-3, 1e-3)
    xm = ops.dropout_forward(x, 0.15, mask)
    ref2 = ops.conv_forward(xm, w, d, S, False, False, None, None, None, False, None, 1.0, ops.QBN_MATH_FP32)
    close(ref, ref2, 1e-5, 1e-5)


S1_SHAPES = [
    # B, C, H, W, N, k
    (2, 24, 32, 32, 24, 3), (2, 48, 16, 16, 48, 3), (3, 96, 8, 8, 96, 3), (5, 192, 4, 4, 192, 3),
    (2, 20, 14, 14, 50, 5), (2, 8, 9, 7, 10, 3), (1, 4, 5, 5, 7, 3), (2, 40, 6, 6, 16, 3), (2, 72, 6, 5, 24, 3),
]


def _tf32_round_(t):
    ti = t.view(torch.int32)
    ti.add_(0x1000).bitwise_and_(~0x1FFF)
    return t


@pytest.mark.parametrize("shape", S1_SHAPES)
def test_tf32_s1_zero_copy_im2col(shape):
    """qbn_conv_s1_fwd (zero-bordered layout, one smem tile for all taps, persistent) == the fp32 conv."""
    from qbn_b200 import ops
    B, C, H, W, N, k = shape
    pad = (k - 1) // 2
    g = torch.Generator().manual_seed(7 + (hash(shape) & 0xFFFF))
    S = 3
    x = _tf32_round_(torch.randn(S * B, C, H, W, generator=g).cuda().contiguous(memory_format=torch.channels_last))
    w = _tf32_round_((torch.randn(S, N, k, k, C, generator=g) / (C * k * k) ** 0.5).cuda()).reshape(S, -1).contiguous()
    scale = (torch.rand(N, generator=g) + 0.5).cuda()
    shift = torch.randn(N, generator=g).cuda()
    res = torch.randn(S * B, N, H, W, generator=g).cuda().contiguous(memory_format=torch.channels_last)
    d = ops.make_desc(B, H, W, C, N, k, k, 1, pad, 1)
    ref = ops.conv_forward(x, w, d, S, False, False, scale, shift, res, True, None, 1.0, ops.QBN_MATH_FP32)
    P = torch.nn.functional.pad
    xp = P(x, (pad, pad, pad, pad)).contiguous(memory_format=torch.channels_last)
    rp = P(res, (pad, pad, pad, pad)).contiguous(memory_format=torch.channels_last)
    got = ops.conv_s1_forward(xp, w, S, N, k, k, scale, shift, rp, True, ops.QBN_FLAG_OUT_ROUND_TF32)
    inner = got[:, :, pad:pad + H, pad:pad + W]
    close(inner, ref, 1e-3, 1e-3)
    border = got.clone()
    border[:, :, pad:pad + H, pad:pad + W] = 0
    assert float(border.abs().max()) == 0.0                              # the zero border is preserved
    assert int((got.view(torch.int32) & 0x1FFF).abs().max()) == 0         # TF32-exact outputs
    # chained: feed the bordered output straight into a second s1 conv (no re-padding)
    w2 = _tf32_round_((torch.randn(S, N, k, k, N, generator=g) / (N * k * k) ** 0.5).cuda()).reshape(S, -1).contiguous()
    if N % 4 == 0:
        got2 = ops.conv_s1_forward(got, w2, S, N, k, k, None, None, None, False, 0)
        d2 = ops.make_desc(B, H, W, N, N, k, k, 1, pad, 1)
        ref2 = ops.conv_forward(inner.contiguous(memory_format=torch.channels_last), w2, d2, S, False, False, None, None, None, False, None, 1.0,
                                ops.QBN_MATH_FP32)
        close(got2[:, :, pad:pad + H, pad:pad + W], ref2, 1e-3, 1e-3)


def test_tf32_v1_bordered_io():
    """v1 gather kernel writing a zero-bordered output, and reading one (stride 2 3x3, 1x1 stride-2 with pad -1)."""
    from qbn_b200 import ops
    g = torch.Generator().manual_seed(11)
    B, C, H, N, S = 2, 24, 16, 48, 2
    x = _tf32_round_(torch.randn(S * B, C, H, H, generator=g).cuda().contiguous(memory_format=torch.channels_last))
    xp = torch.nn.functional.pad(x, (1, 1, 1, 1)).contiguous(memory_format=torch.channels_last)
    for k, stride, pad in ((3, 2, 1), (1, 2, 0), (3, 1, 1)):
        w = _tf32_round_((torch.randn(S, N, k, k, C, generator=g) / (C * k * k) ** 0.5).cuda()).reshape(S, -1).contiguous()
        d = ops.make_desc(B, H, H, C, N, k, k, stride, pad, 1)
        ref = ops.conv_forward(x, w, d, S, False, False, None, None, None, False, None, 1.0, ops.QBN_MATH_FP32)
        dp = ops.make_desc(B, H + 2, H + 2, C, N, k, k, stride, pad - 1, 1)
        assert (dp.Ho, dp.Wo) == (d.Ho, d.Wo)
        dp.out_pad_h = dp.out_pad_w = 1
        out = torch.zeros(S * B, N, d.Ho + 2, d.Wo + 2, device="cuda").contiguous(memory_format=torch.channels_last)
        ops.conv_forward(xp, w, dp, S, False, False, None, None, None, False, None, 1.0, ops.QBN_MATH_TF32, out, ops.QBN_FLAG_A_TF32_READY)
        close(out[:, :, 1:-1, 1:-1], ref, 1e-3, 1e-3)
        b = out.clone()
        b[:, :, 1:-1, 1:-1] = 0
        assert float(b.abs().max()) == 0.0


@pytest.mark.parametrize("shape", [(4, 24, 16, 24, 3, 1, 1), (2, 48, 8, 96, 3, 1, 1), (3, 96, 8, 48, 1, 1, 0), (2, 20, 14, 52, 5, 1, 2),
                                   (2, 24, 16, 48, 3, 2, 1), (8, 192, 1, 12, 1, 1, 0), (2, 24, 16, 48, 1, 2, 0), (2, 8, 7, 12, 3, 2, 1)])
def test_tf32_lrt_backward_dgrad(shape):
    """A3 in TF32 mode against the fp32 kernels: dx on tcgen05 (forward kernel on flipped weights, second launch accumulating
    2x .* conv(dv, sigma2'); transposed-conv gather for strided layers) and the weight gradients on tcgen05 (umma_wgrad.cu)."""
    from qbn_b200 import ops
    B, C, H, N, k, stride, pad = shape
    g = torch.Generator().manual_seed(13 + C + N)
    x = ops.nhwc(torch.randn(B, C, H, H, generator=g).cuda())
    mu = (torch.randn(N, C, k, k, generator=g) / (C * k * k) ** 0.5).cuda()
    rho = (torch.rand(N, C, k, k, generator=g) * 2 - 5).cuda()
    p = ops.weight_prep(mu, rho, False, None, want=("mu", "sigma2"))
    d = ops.make_desc(B, H, H, C, N, k, k, stride, pad, 1)
    out, std = ops.lrt_forward(x, p["mu"], p["sigma2"], None, d, None, (5, 6, 7), ops.QBN_MATH_FP32)
    go = torch.randn(out.shape, generator=g).cuda().contiguous(memory_format=torch.channels_last) if out.dim() == 4 else torch.randn(out.shape, generator=g).cuda()
    ref = ops.lrt_backward(x, p["mu"], p["sigma2"], go, std, d, None, (5, 6, 7), True, False, ops.QBN_MATH_FP32)
    got = ops.lrt_backward(x, p["mu"], p["sigma2"], go, std, d, None, (5, 6, 7), True, False, ops.QBN_MATH_TF32)
    close(got[0], ref[0], 2e-3, 2e-3)
    close(got[1], ref[1], 2e-3, 2e-3)          # dmu: g^T * im2col(x) with the pixels as the tcgen05 reduction dimension
    close(got[2], ref[2], 2e-3, 2e-3)          # dsigma^2: dv^T * im2col(x)^2
