"""GPU (pytest -m gpu): the planar-C4 path — blocked weight sampler, qbn_conv_p4_fwd (stride 1, stride 2 on
phase-split input, phase-split output), planar output of the gather kernel, planar pooling — against the
fp32 CUDA-core convolution on the same TF32-exact operands (rtol 1e-3, north_star's TF32 tolerance)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def close(got, ref, rtol, atol_rel):
    got, ref = got.detach().float().cpu().numpy(), ref.detach().float().cpu().numpy()
    atol = atol_rel * max(1e-30, float(np.abs(ref).max()))
    np.testing.assert_allclose(got, ref, rtol=rtol, atol=atol)


@pytest.fixture(scope="module", autouse=True)
def _lib():
    import __graft_entry__ as g
    g.build()


def _tf32_round_(t):
    ti = t.view(torch.int32)
    ti.add_(0x1000).bitwise_and_(~0x1FFF)
    return t


def torch_conv_fp64(x, w, S, N, k, C, stride, pad, scale=None, shift=None, res=None, relu=False, shared_x=False):
    """Independent reference: torch.nn.functional.conv2d in fp64 on the CPU, sample by sample.  x [S*B (or B), C, H, W], w [S, N*k*k*C]
    packed OHWI (one weight tensor per sample) -> [S*B, N, Ho, Wo].  The TF32-exact operands make the only difference to the tensor-core
    result its fp32 accumulation order and the final TF32 rounding of the output (rtol 1e-3 = north_star's TF32 tolerance)."""
    import torch.nn.functional as F
    xs = x.detach().double().cpu()
    B = xs.shape[0] if shared_x else xs.shape[0] // S
    outs = []
    for s in range(S):
        ws = w[s].detach().double().cpu().reshape(N, k, k, C).permute(0, 3, 1, 2)
        xi = xs if shared_x else xs[s * B:(s + 1) * B]
        y = F.conv2d(xi, ws, None, stride, pad)
        if scale is not None:
            y = y * scale.detach().double().cpu().reshape(1, -1, 1, 1)
        if shift is not None:
            y = y + shift.detach().double().cpu().reshape(1, -1, 1, 1)
        outs.append(y)
    y = torch.cat(outs)
    if res is not None:
        y = y + res.detach().double().cpu()
    return (torch.relu(y) if relu else y).float()


def _rand_weights(g, S, N, k, C):
    return _tf32_round_((torch.randn(S, N, k, k, C, generator=g) / (C * k * k) ** 0.5).cuda()).reshape(S, -1).contiguous()


@pytest.mark.parametrize("shape", [(24, 24, 9, 1), (48, 96, 9, 2), (96, 96, 9, 1), (192, 192, 9, 1), (72, 20, 9, 1), (24, 48, 1, 2), (8, 10, 25, 1),
                                   (96, 192, 9, 2)])
def test_p4_blocked_sampler_matches_canonical(shape):
    """qbn_sample_weights_blocked == qbn_p4_block_weights(qbn_sample_weights): same Philox counters, same roundings."""
    from qbn_b200 import ops
    C, N, taps, stride = shape
    g = torch.Generator().manual_seed(3)
    mu = torch.randn(N * taps * C, generator=g).cuda()
    sg = torch.rand(N * taps * C, generator=g).cuda()
    S = 3
    canon = ops.sample_weights(mu, sg, S, None, 1234, 7, 5, round_tf32=True)
    want = ops.p4_block_weights(canon, N, C, taps, stride)
    mu_b, sg_b = ops.p4_block_weights(mu, N, C, taps, stride)[0], ops.p4_block_weights(sg, N, C, taps, stride)[0]
    got = ops.sample_weights_blocked(mu_b, sg_b, N, C, taps, S, None, 1234, 7, 5, True, stride=stride)
    assert torch.equal(got, want)
    assert got.shape[1] == ops.p4_weight_floats(C, N, taps, 1, stride)
    eps = torch.randn(S, N * taps * C, generator=g).cuda()
    canon = ops.sample_weights(mu, sg, S, eps, 0, 0, 0, round_tf32=False)
    got = ops.sample_weights_blocked(mu_b, sg_b, N, C, taps, S, eps, 0, 0, 0, False, stride=stride)
    assert torch.equal(got, ops.p4_block_weights(canon, N, C, taps, stride))
    # every canonical element appears exactly once, the rest is zero padding
    assert float(got.abs().sum(dtype=torch.float64)) == pytest.approx(float(canon.abs().sum(dtype=torch.float64)), rel=1e-6)


P4_S1 = [
    # B, C, H, W, N, k
    (2, 24, 32, 32, 24, 3), (2, 48, 16, 16, 48, 3), (3, 96, 8, 8, 96, 3), (5, 192, 4, 4, 192, 3),
    (2, 8, 9, 7, 12, 3), (2, 40, 6, 6, 16, 5), (2, 72, 6, 5, 24, 3), (1, 64, 12, 12, 256, 3),
]


@pytest.mark.parametrize("shape", P4_S1)
def test_p4_conv_stride1(shape):
    from qbn_b200 import ops
    B, C, H, W, N, k = shape
    pad = (k - 1) // 2
    g = torch.Generator().manual_seed(11 + C + N)
    S = 3
    x = _tf32_round_(torch.randn(S * B, C, H, W, generator=g).cuda())
    w = _rand_weights(g, S, N, k, C)
    scale = (torch.rand(N, generator=g) + 0.5).cuda()
    shift = torch.randn(N, generator=g).cuda()
    res = torch.randn(S * B, N, H, W, generator=g).cuda()
    d = ops.make_desc(B, H, W, C, N, k, k, 1, pad, 1)
    ref = ops.conv_forward(ops.nhwc(x), w, d, S, False, False, scale, shift, ops.nhwc(res), True, None, 1.0, ops.QBN_MATH_FP32)
    xb = ops.P4Map.from_nchw(x, (pad, pad))
    rb = ops.P4Map.from_nchw(res, (pad, pad))
    wb = ops.p4_block_weights(w, N, C, k * k)
    got = ops.conv_p4_forward(xb, wb, S, N, k, k, 1, scale, shift, rb, True, ops.QBN_FLAG_OUT_ROUND_TF32)
    close(got.to_nchw(), ref, 1e-3, 1e-3)
    ref64 = torch_conv_fp64(x, w, S, N, k, C, 1, pad, scale, shift, res, True)      # not our kernel: torch's conv2d in fp64
    close(got.to_nchw(), ref64, 1e-3, 1e-3)
    close(ref, ref64, 1e-5, 1e-5)                                                   # and the fp32 CUDA-core kernel against it
    full = got.to_nchw(keep_border=True).clone()
    full[:, :, pad:, pad:] = 0
    assert float(full.abs().max()) == 0.0 and float(got.tail().abs().max()) == 0.0       # zero border written, tail untouched
    assert int((got.buf.view(torch.int32) & 0x1FFF).abs().max()) == 0        # TF32-exact outputs
    # chained without re-padding
    w2 = _rand_weights(g, S, N, k, N) if N % 8 == 0 else None
    if w2 is not None:
        got2 = ops.conv_p4_forward(got, ops.p4_block_weights(w2, N, N, k * k), S, N, k, k, 1)
        d2 = ops.make_desc(B, H, W, N, N, k, k, 1, pad, 1)
        ref2 = ops.conv_forward(ops.nhwc(got.to_nchw().contiguous()), w2, d2, S, False, False, None, None, None, False, None, 1.0,
                                ops.QBN_MATH_FP32)
        close(got2.to_nchw(), ref2, 1e-3, 1e-3)
    # phase-split output (for a stride-2 consumer) holds the same values
    if H % 2 == 0 and W % 2 == 0 and k == 3:
        ps = ops.conv_p4_forward(xb, wb, S, N, k, k, 1, scale, shift, rb, True, ops.QBN_FLAG_OUT_ROUND_TF32, phase_split_out=True)
        assert ps.phases == 4
        assert torch.equal(ps.to_nchw(), got.to_nchw())
        rows = ps.buf[:, :4 * S * B * ps.Hp * ps.Wp].permute(1, 0, 2).reshape(4, S * B, ps.Hp, ps.Wp, N).clone()
        rows[:, :, 1:, 1:, :] = 0
        assert float(rows.abs().max()) == 0.0


@pytest.mark.parametrize("shape", [(2, 24, 32, 48, 3), (2, 24, 32, 48, 1), (3, 48, 16, 96, 3), (3, 48, 16, 96, 1), (2, 96, 8, 192, 3),
                                   (2, 96, 8, 192, 1), (1, 8, 6, 12, 3)])
def test_p4_conv_stride2_phase_split(shape):
    from qbn_b200 import ops
    B, C, H, N, k = shape
    pad = (k - 1) // 2
    g = torch.Generator().manual_seed(5 + C + k)
    S = 2
    x = _tf32_round_(torch.randn(S * B, C, H, H, generator=g).cuda())
    w = _rand_weights(g, S, N, k, C)
    scale = (torch.rand(N, generator=g) + 0.5).cuda()
    shift = torch.randn(N, generator=g).cuda()
    d = ops.make_desc(B, H, H, C, N, k, k, 2, pad, 1)
    ref = ops.conv_forward(ops.nhwc(x), w, d, S, False, False, scale, shift, None, False, None, 1.0, ops.QBN_MATH_FP32)
    xs = ops.P4Map.from_nchw(x, None, phase_split=True)
    assert torch.equal(xs.to_nchw(), x)
    got = ops.conv_p4_forward(xs, ops.p4_block_weights(w, N, C, k * k, 2), S, N, k, k, 2, scale, shift, None, False, ops.QBN_FLAG_OUT_ROUND_TF32)
    assert (got.Hp, got.Wp) == (H // 2 + 1, H // 2 + 1)
    close(got.to_nchw(), ref, 1e-3, 1e-3)
    close(got.to_nchw(), torch_conv_fp64(x, w, S, N, k, C, 2, pad, scale, shift), 1e-3, 1e-3)
    full = got.to_nchw(keep_border=True).clone()
    full[:, :, 1:, 1:] = 0
    assert float(full.abs().max()) == 0.0


def test_v1_planar_output_and_pool():
    """The gather kernel (first layer: shared input, sample-stacked weights) writing planar C4 directly,
    and the planar global average pool."""
    from qbn_b200 import ops
    g = torch.Generator().manual_seed(21)
    B, S, N = 3, 4, 24
    x = torch.randn(B, 4, 16, 16, generator=g).cuda()
    w = _rand_weights(g, S, N, 3, 4)
    scale = (torch.rand(N, generator=g) + 0.5).cuda()
    shift = torch.randn(N, generator=g).cuda()
    d = ops.make_desc(B, 16, 16, 4, N, 3, 3, 1, 1, 1)
    ref = ops.conv_forward(ops.nhwc(x), w, d, S, True, False, scale, shift, None, True, None, 1.0, ops.QBN_MATH_TF32)
    d.out_pad_h = d.out_pad_w = 1
    out = ops.P4Map.empty(S * B, N, 17, 17, (1, 1), 1, "cuda")
    ops.conv_forward(ops.nhwc(x), w, d, S, True, False, scale, shift, None, True, None, 1.0, ops.QBN_MATH_TF32, out.buf, ops.QBN_FLAG_OUT_P4)
    assert torch.equal(out.to_nchw(), ref.reshape(S * B, N, 16, 16))
    # per-sample (not stacked) launch with a planar residual
    res = ops.P4Map.from_nchw(torch.randn(S * B, N, 16, 16, generator=g).cuda(), (1, 1))
    xs = _tf32_round_(torch.randn(S * B, 8, 16, 16, generator=g).cuda())
    w8 = _rand_weights(g, S, N, 3, 8)
    d8 = ops.make_desc(B, 16, 16, 8, N, 3, 3, 1, 1, 1)
    ref = ops.conv_forward(ops.nhwc(xs), w8, d8, S, False, False, scale, shift, ops.nhwc(res.to_nchw().contiguous()), True, None, 1.0,
                           ops.QBN_MATH_TF32)
    d8.out_pad_h = d8.out_pad_w = 1
    out2 = ops.P4Map.empty(S * B, N, 17, 17, (1, 1), 1, "cuda")
    ops.conv_forward(ops.nhwc(xs), w8, d8, S, False, False, scale, shift, res.buf, True, None, 1.0, ops.QBN_MATH_TF32, out2.buf,
                     ops.QBN_FLAG_OUT_P4)
    assert torch.equal(out2.to_nchw(), ref)
    pooled = ops.avgpool_p4(out2, 16 * 16)
    close(pooled, ref.mean(dim=(2, 3)), 1e-5, 1e-6)


def test_p4_conv_full_size_linearity():
    """BASELINE-size tile counts (B=256, 10 samples, 23 120 tiles): conv(x, w1 + w2) == conv(x, w1) + conv(x, w2) up to
    TF32 operand rounding, and a second run is bit-identical (no race in the persistent pipeline)."""
    from qbn_b200 import ops
    g = torch.Generator().manual_seed(2)
    S, B, C, H = 10, 256, 24, 32
    x = _tf32_round_(torch.randn(S * B, C, H, H, generator=g).cuda())
    xb = ops.P4Map.from_nchw(x, (1, 1))
    del x
    w1 = _rand_weights(g, S, C, 3, C)
    w2 = _rand_weights(g, S, C, 3, C)
    w12 = _tf32_round_((w1 + w2).clone())
    blk = lambda w: ops.p4_block_weights(w, C, C, 9)
    a = ops.conv_p4_forward(xb, blk(w1), S, C, 3, 3, 1).buf
    b = ops.conv_p4_forward(xb, blk(w2), S, C, 3, 3, 1).buf
    c = ops.conv_p4_forward(xb, blk(w12), S, C, 3, 3, 1).buf
    c2 = ops.conv_p4_forward(xb, blk(w12), S, C, 3, 3, 1).buf
    assert torch.equal(c, c2)
    err = float((a + b - c).abs().max()) / float(c.abs().max())
    assert err < 2e-3, err


def _stack_blocked(wb, N, n_pad, S):
    """[S, cols*n_pad*4] per-sample blocked weights -> ONE blocked tensor whose rows are the samples stacked."""
    cols = wb.shape[1] // (n_pad * 4)
    rows = wb.reshape(S, cols, n_pad, 4)[:, :, :N, :]                 # [S, cols, N, 4]
    st = rows.permute(1, 0, 2, 3).reshape(cols, S * N, 4)
    n_pad_st = (S * N + 15) // 16 * 16
    out = torch.zeros(cols, n_pad_st, 4, device=wb.device)
    out[:, :S * N] = st
    return out.reshape(1, -1).contiguous()


def test_p4_first_layer_sample_stacked():
    """Shared input + the samples' weights stacked along N (QBN_FLAG_X_SHARED_STACKED) == per-sample convolutions."""
    from qbn_b200 import ops
    g = torch.Generator().manual_seed(31)
    B, S, C, N, H = 3, 4, 8, 24, 16
    x = _tf32_round_(torch.randn(B, C, H, H, generator=g).cuda())
    x[:, 3:] = 0
    w = _rand_weights(g, S, N, 3, C)
    scale = (torch.rand(N, generator=g) + 0.5).cuda()
    shift = torch.randn(N, generator=g).cuda()
    d = ops.make_desc(B, H, H, C, N, 3, 3, 1, 1, 1)
    ref = ops.conv_forward(ops.nhwc(x), w, d, S, True, False, scale, shift, None, True, None, 1.0, ops.QBN_MATH_FP32).reshape(S * B, N, H, H)
    wb = ops.p4_block_weights(w, N, C, 9)
    wst = _stack_blocked(wb, N, 32, S)
    assert wst.shape[1] == ops.p4_weight_floats(C, S * N, 3, 3, 1)
    xm = ops.P4Map.from_nchw(x, (1, 1))
    got = ops.conv_p4_forward(xm, wst, S, N, 3, 3, 1, scale, shift, None, True, ops.QBN_FLAG_OUT_ROUND_TF32 | ops.QBN_FLAG_X_SHARED_STACKED)
    assert got.n_img == S * B
    close(got.to_nchw(), ref, 1e-3, 1e-3)
    close(got.to_nchw(), torch_conv_fp64(x, w, S, N, 3, C, 1, 1, scale, shift, None, True, shared_x=True), 1e-3, 1e-3)
    full = got.to_nchw(keep_border=True).clone()
    full[:, :, 1:, 1:] = 0
    assert float(full.abs().max()) == 0.0


def test_p4_multi_layer_sampler():
    """qbn_sample_weights_blocked_multi (all layers of a chunk in one launch, incl. the stacked first layer) is
    bit-identical to the per-layer sampler."""
    from qbn_b200 import ops
    from qbn_b200._lib import P4SampleJob
    g = torch.Generator().manual_seed(41)
    S = 5
    layers = [(8, 24, 9, 1, 11, True), (24, 24, 9, 1, 12, False), (24, 48, 9, 2, 13, False), (24, 48, 1, 2, 14, False), (96, 96, 9, 1, 15, False)]
    jobs = (P4SampleJob * len(layers))()
    keep, want = [], []
    for i, (C, N, taps, stride, lid, stack) in enumerate(layers):
        mu = torch.randn(N * taps * C, generator=g).cuda()
        sg = torch.rand(N * taps * C, generator=g).cuda()
        mu_b, sg_b = ops.p4_block_weights(mu, N, C, taps, stride)[0], ops.p4_block_weights(sg, N, C, taps, stride)[0]
        single = ops.sample_weights_blocked(mu_b, sg_b, N, C, taps, S, None, 77, lid, 3, True, None, stride)
        if stack:
            single = _stack_blocked(single, N, (N + 15) // 16 * 16, S)
            w = torch.zeros_like(single)
        else:
            w = torch.empty_like(single)
        jobs[i] = P4SampleJob(mu_b.data_ptr(), sg_b.data_ptr(), None, w.data_ptr(), N, C, taps, stride, lid, S if stack else 0)
        keep.append((mu_b, sg_b, w))
        want.append(single)
    raw = torch.frombuffer(bytearray(bytes(jobs)), dtype=torch.uint8).cuda()
    ops.sample_weights_blocked_multi(raw, len(layers), max(k[0].numel() for k in keep), S, 77, 3, True)
    for (mu_b, sg_b, w), ref in zip(keep, want):
        assert torch.equal(w, ref)


def test_p4_conv_output_dropout_mask():
    """Epilogue order affine -> ReLU(pre) -> mask*mult -> +residual -> ReLU(post): the MC-Dropout BasicBlock of models_mc.py."""
    from qbn_b200 import ops
    g = torch.Generator().manual_seed(91)
    S, B, C, N, H = 2, 3, 24, 24, 8
    x = _tf32_round_(torch.randn(S * B, C, H, H, generator=g).cuda())
    w = _rand_weights(g, 1, N, 3, C)
    scale = (torch.rand(N, generator=g) + 0.5).cuda()
    shift = torch.randn(N, generator=g).cuda()
    res = torch.randn(S * B, N, H, H, generator=g).cuda()
    mask = (torch.rand(S * B, N, generator=g) < 0.8).float().cuda()
    mult = 1.0 / 0.85
    conv = torch.nn.functional.conv2d(x.double(), w.reshape(N, 3, 3, C).permute(0, 3, 1, 2).double(), None, 1, 1)
    aff = conv * scale.double().view(1, -1, 1, 1) + shift.double().view(1, -1, 1, 1)
    m4 = mask.double().view(S * B, N, 1, 1)
    xb, rb, wb = ops.P4Map.from_nchw(x, (1, 1)), ops.P4Map.from_nchw(res, (1, 1)), ops.p4_block_weights(w, N, C, 9)
    # conv-BN-ReLU-dropout
    got = ops.conv_p4_forward(xb, wb, S, N, 3, 3, 1, scale, shift, None, False, ops.QBN_FLAG_RELU_PRE, True, None, False, mask, mult)
    close(got.to_nchw(), (aff.clamp(min=0) * m4 * mult).float(), 1e-3, 1e-3)
    # conv-BN-dropout-add-ReLU
    got = ops.conv_p4_forward(xb, wb, S, N, 3, 3, 1, scale, shift, rb, True, 0, True, None, False, mask, mult)
    close(got.to_nchw(), ((aff * m4 * mult + res.double()).clamp(min=0)).float(), 1e-3, 1e-3)
    # first layer, shared input, per-sample masks on identical (stacked) weights
    x1 = _tf32_round_(torch.randn(B, 8, H, H, generator=g).cuda())
    w1 = _rand_weights(g, 1, N, 3, 8)
    wst = _stack_blocked(ops.p4_block_weights(w1, N, 8, 9).repeat(S, 1), N, 32, S)
    got = ops.conv_p4_forward(ops.P4Map.from_nchw(x1, (1, 1)), wst, S, N, 3, 3, 1, scale, shift, None, False,
                              ops.QBN_FLAG_RELU_PRE | ops.QBN_FLAG_X_SHARED_STACKED, False, None, False, mask, mult)
    conv1 = torch.nn.functional.conv2d(x1.double(), w1.reshape(N, 3, 3, 8).permute(0, 3, 1, 2).double(), None, 1, 1)
    aff1 = (conv1 * scale.double().view(1, -1, 1, 1) + shift.double().view(1, -1, 1, 1)).clamp(min=0)
    ref = (aff1.unsqueeze(0) * mask.double().view(S, B, N, 1, 1) * mult).reshape(S * B, N, H, H)
    close(got.to_nchw(), ref.float(), 1e-3, 1e-3)


@pytest.mark.parametrize("shape", [(2, 24, 48, 16), (2, 48, 96, 8), (3, 96, 192, 4)])
def test_p4_conv_fused_shortcut(shape):
    """conv3x3(y) * s2 + conv1x1,stride2(x) * ssc + shift in ONE accumulator (scales folded into the weights) ==
    BN2(conv2(y)) + BNsc(convsc(x)) of models_bbb.py:170-178."""
    from qbn_b200 import ops
    B, C2, N, Ho = shape          # block input x: C2 channels at 2*Ho; main conv: N -> N at Ho
    g = torch.Generator().manual_seed(100 + N)
    S = 3
    x = _tf32_round_(torch.randn(S * B, C2, 2 * Ho, 2 * Ho, generator=g).cuda())
    y = _tf32_round_(torch.randn(S * B, N, Ho, Ho, generator=g).cuda())
    w = _rand_weights(g, S, N, 3, N)
    wsc = _rand_weights(g, S, N, 1, C2)
    s2, ssc = (torch.rand(N, generator=g) + 0.5).cuda(), (torch.rand(N, generator=g) + 0.5).cuda()
    shift = torch.randn(N, generator=g).cuda()
    d = ops.make_desc(B, Ho, Ho, N, N, 3, 3, 1, 1, 1)
    dsc = ops.make_desc(B, 2 * Ho, 2 * Ho, C2, N, 1, 1, 2, 0, 1)
    sc = ops.conv_forward(ops.nhwc(x), wsc, dsc, S, False, False, ssc, None, None, False, None, 1.0, ops.QBN_MATH_FP32)
    ref = ops.conv_forward(ops.nhwc(y), w, d, S, False, False, s2, shift, sc, True, None, 1.0, ops.QBN_MATH_FP32)
    # fold the scales into the (TF32-rounded) weights like the sampler does
    wf = _tf32_round_((w.reshape(S, N, -1) * s2.view(1, N, 1)).reshape(S, -1).contiguous())
    wscf = _tf32_round_((wsc.reshape(S, N, -1) * ssc.view(1, N, 1)).reshape(S, -1).contiguous())
    cb2 = ops.p4_shortcut_block_channels(N, C2)
    assert cb2 % 8 == 0 and C2 % cb2 == 0
    wb = torch.cat([ops.p4_block_weights(wf, N, N, 9), ops.p4_block_weights(wscf, N, C2, 1, 2, None, cb2)], dim=1).contiguous()
    yb = ops.P4Map.from_nchw(y, (1, 1))
    xs = ops.P4Map.from_nchw(x, None, phase_split=True)
    got = ops.conv_p4_shortcut_forward(yb, wb, xs, S, N, 3, 3, None, shift, True, ops.QBN_FLAG_OUT_ROUND_TF32)
    close(got.to_nchw(), ref, 2e-3, 2e-3)
    # independent fp64 reference: relu(conv3x3(y) * s2 + conv1x1/2(x) * ssc + shift)  (2e-3: the scales are folded into re-rounded weights)
    sc64 = torch_conv_fp64(x, wsc, S, N, 1, C2, 2, 0, ssc)
    close(got.to_nchw(), torch_conv_fp64(y, w, S, N, 3, N, 1, 1, s2, shift, sc64, True), 2e-3, 2e-3)
    full = got.to_nchw(keep_border=True).clone()
    full[:, :, 1:, 1:] = 0
    assert float(full.abs().max()) == 0.0


def test_avgpool_p4_small_maps():
    """qbn_avgpool_p4 on the 5x5 padded map that ends the ResNet (one thread per (image, chunk)) and on a larger map (one warp each)."""
    from qbn_b200 import ops
    g = torch.Generator().manual_seed(6)
    for (n, C, H) in ((37, 192, 4), (5, 24, 7), (3, 48, 16)):
        x = torch.randn(n, C, H, H, generator=g).cuda()
        m = ops.P4Map.from_nchw(x, (1, 1))
        pooled = ops.avgpool_p4(m, H * H)
        close(pooled, x.mean(dim=(2, 3)), 1e-5, 1e-6)
