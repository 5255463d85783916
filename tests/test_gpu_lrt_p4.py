"""GPU (pytest -m gpu): LRT training on the planar zero-copy tcgen05 kernels — staging, forward, input
gradient (stride 1 and the phase launches of stride 2), weight gradients — against the oracle (bbb/conv.py:24-32 and its
autograd).  TF32 tolerance per north_star: rtol 1e-3, atol 1e-3 * max|ref| (10 mantissa bits per operand, fp32 accumulation)."""
import numpy as np
import pytest
import torch

import oracle.qbn_oracle as O

pytestmark = pytest.mark.gpu

SHAPES = [
    # B, C, H, N, k, stride, pad
    (2, 24, 32, 24, 3, 1, 1),     # ResNet layer1
    (2, 3, 32, 24, 3, 1, 1),      # first layer: 3 -> 8 padded input channels (no input gradient)
    (2, 24, 32, 48, 3, 2, 1),     # layer2 stride-2 entry
    (2, 24, 32, 48, 1, 2, 0),     # 1x1 stride-2 shortcut
    (2, 48, 16, 48, 3, 1, 1),
    (3, 96, 8, 96, 3, 1, 1),
    (3, 96, 8, 192, 3, 2, 1),
    (3, 96, 8, 192, 1, 2, 0),
    (5, 192, 4, 192, 3, 1, 1),    # 125 padded pixels: less than one tile
    (2, 16, 12, 40, 5, 1, 2),     # 5x5 'same'
    (3, 8, 9, 16, 3, 1, 1),       # odd sizes
    (4, 32, 6, 8, 1, 1, 0),       # 1x1 stride 1 (no border)
    (33, 24, 16, 24, 3, 1, 1),    # several tiles per CTA, ragged last tile
]


def close(got, ref, rtol=1e-3, atol_rel=1e-3):
    got, ref = got.detach().float().cpu().numpy(), ref.detach().float().cpu().numpy()
    atol = atol_rel * max(1e-30, float(np.abs(ref).max()))
    np.testing.assert_allclose(got, ref, rtol=rtol, atol=atol)


@pytest.fixture(scope="module", autouse=True)
def _lib():
    import __graft_entry__ as g
    g.build()


def _case(shape, seed=0):
    B, C, H, N, k, stride, pad = shape
    g = torch.Generator().manual_seed(seed + (hash(shape) & 0xFFFF))
    x = torch.randn(B, C, H, H, generator=g)
    mu = torch.randn(N, C, k, k, generator=g) / (C * k * k) ** 0.5
    rho = torch.empty(N, C, k, k).uniform_(-5, -2, generator=g)
    Ho = (H + 2 * pad - k) // stride + 1
    eps = torch.randn(B, N, Ho, Ho, generator=g)
    go = torch.randn(B, N, Ho, Ho, generator=g)
    return x, mu, rho, eps, go


def test_stage_input_layout():
    """qbn_p4_stage_input against the torch restatement of the planar layout (ops.P4Map.from_nchw), normal and phase-split."""
    from qbn_b200 import ops
    g = torch.Generator().manual_seed(3)
    for (B, C, H, W, border, split) in ((3, 24, 8, 6, (1, 1), False), (2, 3, 6, 6, (1, 1), False), (2, 16, 8, 12, (1, 1), True),
                                        (2, 8, 5, 7, (2, 2), False), (2, 8, 4, 4, (0, 0), False)):
        x = torch.randn(B, C, H, W, generator=g).cuda()
        xc = ops.nhwc(x)
        C_pad = (C + 7) // 8 * 8
        Hp, Wp = (H // 2 + 1, W // 2 + 1) if split else (H + border[0], W + border[1])
        rows = (4 if split else 1) * B * Hp * Wp
        pr = ops.lrt_p4_plane_rows(rows, Wp)
        xp = torch.full((C_pad // 4, pr, 4), float("nan"), device="cuda")
        xq = torch.full((C_pad // 4, pr, 4), float("nan"), device="cuda")
        ops._lib.call("qbn_p4_stage_input", ops._ptr(xc), B, H, W, C, C_pad, border[0], border[1], int(split), pr, ops._ptr(xp), ops._ptr(xq),
                      ops._stream())
        xpad = torch.nn.functional.pad(x, (0, 0, 0, 0, 0, C_pad - C))
        ref = ops.P4Map.from_nchw(xpad, border, phase_split=split).buf
        assert torch.isfinite(xp).all() and torch.isfinite(xq).all()
        np.testing.assert_allclose(xp[:, :ref.shape[1]].cpu().numpy(), ref.cpu().numpy(), rtol=6e-4, atol=0)       # RNA to 10 mantissa bits
        np.testing.assert_allclose(xq[:, :ref.shape[1]].cpu().numpy(), (ref * ref).cpu().numpy(), rtol=6e-4, atol=0)
        assert float(xp[:, ref.shape[1]:].abs().max()) == 0.0 and float(xq[:, rows:].abs().max()) == 0.0           # zero tail
        assert int((xp.view(torch.int32) & 0x1FFF).abs().max()) == 0                                               # TF32-exact


def test_w32_layout():
    """qbn_w32_from_p4: [C/4][rows][4] -> [ceil(C/32)][rows][32] with the 32-byte chunks of row r at position chunk ^ (r & 3) — the global
    image of SWIZZLE_128B_BASE32B (the only shared-memory layout tcgen05.mma kind::tf32 takes for MN-major operands)."""
    from qbn_b200 import ops
    g = torch.Generator().manual_seed(9)
    for chunks, rows in ((6, 37), (2, 8), (12, 130), (24, 21), (48, 9)):
        a = torch.randn(chunks, rows, 4, generator=g).cuda()
        b = torch.randn(chunks, rows, 4, generator=g).cuda()
        wa, wb = ops.w32_from_p4(a, b)
        wa1 = ops.w32_from_p4(a)
        nblk = (chunks + 7) // 8
        for src, got in ((a, wa), (b, wb), (a, wa1)):
            dense = torch.zeros(nblk * 8, rows, 4, device="cuda")
            dense[:chunks] = src
            ref = dense.view(nblk, 4, 2, rows, 4).permute(0, 3, 1, 2, 4).reshape(nblk, rows, 4, 8)       # [blk][row][chunk][8 channels]
            r = torch.arange(rows, device="cuda") & 3
            pos = torch.arange(4, device="cuda")[None, :] ^ r[:, None]                                   # chunk c of row r sits at c ^ (r & 3)
            exp = torch.zeros_like(ref)
            exp.scatter_(2, pos[None, :, :, None].expand(nblk, rows, 4, 8), ref)
            assert torch.equal(got.view(nblk, rows, 4, 8), exp)


def test_fused_staging_equals_the_two_pass_staging():
    """qbn_lrt_stage_input / _grad (one pass, both layouts) == qbn_p4_stage_* followed by qbn_w32_from_p4, bit for bit."""
    from qbn_b200 import ops
    g = torch.Generator().manual_seed(4)
    for (B, C, H, W, border, split) in ((3, 24, 8, 6, (1, 1), False), (2, 3, 6, 6, (1, 1), False), (2, 16, 8, 12, (1, 1), True),
                                        (2, 40, 5, 7, (2, 2), False), (5, 96, 4, 4, (0, 0), False)):
        x = ops.nhwc(torch.randn(B, C, H, W, generator=g).cuda())
        C_pad = (C + 7) // 8 * 8
        Hp, Wp = (H // 2 + 1, W // 2 + 1) if split else (H + border[0], W + border[1])
        pr = ops.lrt_p4_plane_rows((4 if split else 1) * B * Hp * Wp, Wp)
        mk = lambda *shape: torch.full(shape, float("nan"), device="cuda")
        a, b = mk(C_pad // 4, pr, 4), mk(C_pad // 4, pr, 4)
        ops._lib.call("qbn_p4_stage_input", ops._ptr(x), B, H, W, C, C_pad, border[0], border[1], int(split), pr, ops._ptr(a), ops._ptr(b), ops._stream())
        wa, wb = ops.w32_from_p4(a, b)
        a2, b2, wa2, wb2 = mk(C_pad // 4, pr, 4), mk(C_pad // 4, pr, 4), mk((C_pad + 31) // 32, pr, 32), mk((C_pad + 31) // 32, pr, 32)
        ops._lib.call("qbn_lrt_stage_input", ops._ptr(x), B, H, W, C, C_pad, border[0], border[1], int(split), pr, ops._ptr(a2), ops._ptr(b2),
                      ops._ptr(wa2), ops._ptr(wb2), ops._stream())
        for u, v in ((a, a2), (b, b2), (wa, wa2), (wb, wb2)):
            assert torch.equal(u, v)
        # ... and with the layer's noise tensor drawn in the same launch: same layouts, the values of qbn_lrt_noise (more and fewer
        # noise elements than staging work; a device-side draw offset)
        from qbn_b200 import noise
        for n_noise, off in ((B * H * W * 8, 0), (4 * 1000 * 1000, 3)):
            noise.set_draw_offset(off)
            want = mk(n_noise)
            ops._lib.call("qbn_lrt_noise", ops._ptr(want), n_noise, 11, 12, 13, ops._stream())
            a3, b3, wa3, wb3, got = mk(C_pad // 4, pr, 4), mk(C_pad // 4, pr, 4), mk((C_pad + 31) // 32, pr, 32), mk((C_pad + 31) // 32, pr, 32), mk(n_noise)
            ops._lib.call("qbn_lrt_stage_input_noise", ops._ptr(x), B, H, W, C, C_pad, border[0], border[1], int(split), pr, ops._ptr(a3), ops._ptr(b3),
                          ops._ptr(wa3), ops._ptr(wb3), ops._ptr(got), n_noise, 11, 12, 13, ops._stream())
            for u, v in ((a, a3), (b, b3), (wa, wa3), (wb, wb3), (want, got)):
                assert torch.equal(u, v)
        noise.set_draw_offset(0)
        if not split and C % 4 == 0:
            gout, sd = ops.nhwc(torch.randn(B, C, H, W, generator=g).cuda()), ops.nhwc(torch.rand(B, C, H, W, generator=g).cuda() + 0.5)
            for eps in (ops.nhwc(torch.randn(B, C, H, W, generator=g).cuda()), None):
                a, b = mk(C // 4, pr, 4), mk(C // 4, pr, 4)
                ops._lib.call("qbn_p4_stage_grad", ops._ptr(gout), ops._ptr(sd), ops._ptr(eps), 5, 6, 7, B, H, W, C, border[0], border[1], pr, ops._ptr(a),
                              ops._ptr(b), ops._stream())
                wa, wb = ops.w32_from_p4(a, b)
                a2, b2, wa2, wb2 = mk(C // 4, pr, 4), mk(C // 4, pr, 4), mk((C + 31) // 32, pr, 32), mk((C + 31) // 32, pr, 32)
                ops._lib.call("qbn_lrt_stage_grad", ops._ptr(gout), ops._ptr(sd), ops._ptr(eps), 5, 6, 7, B, H, W, C, border[0], border[1], pr, ops._ptr(a2),
                              ops._ptr(b2), ops._ptr(wa2), ops._ptr(wb2), ops._stream())
                for u, v in ((a, a2), (b, b2), (wa, wa2), (wb, wb2)):
                    assert torch.equal(u, v)


@pytest.mark.parametrize("shape", SHAPES)
def test_lrt_p4_forward(shape):
    from qbn_b200 import ops
    B, C, H, N, k, stride, pad = shape
    x, mu, rho, eps, _ = _case(shape)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(5))
    d = ops.make_desc(B, H, H, C, N, k, k, stride, pad, 1)
    assert ops.lrt_p4_eligible(d, need_dx=C % 8 == 0)
    out, std, _, _ = ops.lrt_p4_forward(ops.nhwc(x.cuda()), mu.cuda(), rho.cuda(), False, bias.cuda(), d, ops.nhwc(eps.cuda()), want_w32=False)
    yo, so = O.lrt_conv_fwd(x, mu, rho, bias, eps, stride, pad)
    close(std, so)
    close(out, yo)
    # Philox epilogue: the draw of the gather kernel (counter = offset in out / 4)
    p = ops.weight_prep(mu.cuda(), rho.cuda(), want=("mu", "sigma2"))
    r1, _ = ops.lrt_forward(ops.nhwc(x.cuda()), p["mu"], p["sigma2"], None, d, None, (5, 6, 7), ops.QBN_MATH_FP32)
    r2, _, _, _ = ops.lrt_p4_forward(ops.nhwc(x.cuda()), mu.cuda(), rho.cuda(), False, None, d, None, (5, 6, 7))
    close(r2, r1, 1e-3, 2e-3)


@pytest.mark.parametrize("shape", SHAPES)
def test_lrt_p4_backward(shape):
    from qbn_b200 import ops
    B, C, H, N, k, stride, pad = shape
    x, mu, rho, eps, go = _case(shape, 7)
    d = ops.make_desc(B, H, H, C, N, k, k, stride, pad, 1)
    need_dx = C % 8 == 0
    xc = ops.nhwc(x.cuda())
    out, std, x_p4, xsq_p4 = ops.lrt_p4_forward(xc, mu.cuda(), rho.cuda(), False, None, d, ops.nhwc(eps.cuda()))
    dx, dmu_p, dsig2_p = ops.lrt_p4_backward(xc, x_p4, xsq_p4, std, ops.nhwc(eps.cuda()), mu.cuda(), rho.cuda(), False, ops.nhwc(go.cuda()), d,
                                             (0, 0, 0), need_dx)
    d_mu, d_rho = ops.weight_grad_post(dmu_p, dsig2_p, rho.cuda(), False, tuple(mu.shape))
    _, so = O.lrt_conv_fwd(x, mu, rho, None, eps, stride, pad)
    dx_o, dmu_o, drho_o, _ = O.lrt_conv_bwd(x, mu, rho, eps, so, go, stride, pad)
    problems = []
    for name, got, ref in (("dx", dx, dx_o), ("d_mu", d_mu, dmu_o), ("d_rho", d_rho, drho_o)):
        if name == "dx" and not need_dx:
            assert dx is None
            continue
        try:
            close(got, ref)
        except AssertionError as e:
            gn, rn = got.detach().float().cpu(), ref.detach().float().cpu()
            problems.append("%s: max|got| %.4g max|ref| %.4g max|diff| %.4g\n%s" % (name, float(gn.abs().max()), float(rn.abs().max()),
                                                                                 float((gn - rn).abs().max()), str(e)[:300]))
    assert not problems, "\n".join(problems)


@pytest.mark.parametrize("shape", [SHAPES[0], SHAPES[2], SHAPES[3], SHAPES[5]])
def test_lrt_function_takes_the_planar_path_and_matches_the_fp32_mode(shape):
    """ops.LRTFunction (what stochastic.bbb.Conv2d.forward calls in training mode): TF32 mode runs planar, same Philox draw as fp32 mode."""
    from qbn_b200 import ops
    B, C, H, N, k, stride, pad = shape
    x, mu, rho, _, go = _case(shape, 11)
    res = {}
    for mode in (ops.QBN_MATH_FP32, ops.QBN_MATH_TF32):
        xx = x.cuda().requires_grad_(True)
        m, r = mu.cuda().requires_grad_(True), rho.cuda().requires_grad_(True)
        calls = []
        real = ops._lib.call
        ops._lib.call = lambda name, *a: (calls.append(name), real(name, *a))[1]
        try:
            out = ops.LRTFunction.apply(xx, m, r, None, (stride, stride), (pad, pad), (1, 1), None, (9, 3, 1), mode, False, None)
            (out * go.cuda()).sum().backward()
        finally:
            ops._lib.call = real
        assert ("qbn_lrt_conv_p4_fwd" in calls) == (mode == ops.QBN_MATH_TF32)
        assert ("qbn_lrt_wgrad_p4" in calls) == (mode == ops.QBN_MATH_TF32)
        res[mode] = (out.detach(), xx.grad, m.grad, r.grad)
    for a, b in zip(res[ops.QBN_MATH_TF32], res[ops.QBN_MATH_FP32]):
        close(a, b, 1e-3, 2e-3)


def test_lrt_p4_ineligible_shapes_stay_on_the_gather_kernels():
    from qbn_b200 import ops
    for (C, H, N, k, stride, pad, dil) in ((20, 14, 50, 5, 1, 2, 1), (24, 16, 24, 3, 1, 0, 1), (24, 16, 24, 3, 1, 2, 2), (24, 16, 20, 3, 1, 1, 1),
                                           (24, 15, 24, 3, 2, 1, 1), (24, 16, 264, 3, 1, 1, 1)):
        assert not ops.lrt_p4_eligible(ops.make_desc(2, H, H, C, N, k, k, stride, pad, dil))
    assert not ops.lrt_p4_eligible(ops.make_desc(2, 32, 32, 3, 24, 3, 3, 1, 1, 1), need_dx=True)
    assert ops.lrt_p4_eligible(ops.make_desc(2, 32, 32, 3, 24, 3, 3, 1, 1, 1), need_dx=False)


def test_graphed_train_step_matches_the_eager_loop():
    """dist.GraphedTrainStep (one CUDA graph per step, device-side draw offset) == DPTrainStep run eagerly for the same number of
    steps: same noise sequence, same parameters up to the summation order of the weight-gradient atomics."""
    from qbn_b200 import config, losses, noise, synthetic, zoo
    from qbn_b200 import dist as qdist
    config.set_math_mode("tf32")
    g = torch.Generator().manual_seed(5)
    x, t = torch.randn(16, 3, 32, 32, generator=g).cuda(), torch.randint(0, 10, (16,), generator=g).cuda()
    crit = losses.LOSS_FACTORY["classification"](zoo.Args(loss_multiplier=1.0), "batch")
    finals, losses_seen = [], []
    first_id = noise._state["next_layer_id"]
    for graphed in (False, True):
        noise._state["next_layer_id"] = first_id          # the same Philox stream ids (layer ids) for both models
        model = zoo.resnet_from_params(synthetic.ResNetBBBParams(seed=1)).cuda().train()
        noise.manual_seed(77)
        noise.set_draw_offset(0)
        # plain momentum SGD: the update is linear in the gradient, so the summation order of the weight-gradient atomics stays a
        # rounding-level difference (Adam's early steps are sign-like and amplify it to the size of the step itself)
        opt = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=1e-3, momentum=0.9)
        if graphed:
            step = qdist.GraphedTrainStep(model, crit, opt, x, t, 176, 45000, gamma=0.01, warmup=2)
            for _ in range(3):
                out = step(x, t)
            assert step.draws_per_step == 21 and step.steps_done == 5
        else:
            step = qdist.DPTrainStep(model, crit, opt, gamma=0.01, check_nan_loss=False)
            for _ in range(5):
                out = step(x, t, 176, 45000)
        torch.cuda.synchronize()
        losses_seen.append(float(out[1]))
        finals.append({k: v.detach().clone() for k, v in model.named_parameters()})
        noise.set_draw_offset(0)
    assert abs(losses_seen[0] - losses_seen[1]) < 2e-3 * abs(losses_seen[0])
    for k in finals[0]:
        a, b = finals[0][k], finals[1][k]
        assert float((a - b).norm() / (a.norm() + 1e-12)) < 2e-3, k
    config.set_math_mode("fp32")


def test_nan_scrub_of_all_gradients_in_one_launch():
    """dist.scrub_nan_grads == trainer.py:105-107 (`p.grad[p.grad != p.grad] = 0` per parameter; infinities stay)."""
    from qbn_b200 import dist as qdist
    g = torch.Generator().manual_seed(8)
    params = [torch.nn.Parameter(torch.randn(n, generator=g).cuda()) for n in (1, 7, 1024, 5000, 3)] + [torch.nn.Parameter(torch.randn(4).cuda())]
    params[-1].grad = None
    refs = []
    for p in params[:-1]:
        gr = torch.randn(p.shape, generator=g).cuda()
        gr[::3] = float("nan")
        gr[1::7] = float("inf")
        p.grad = gr
        r = gr.clone()
        r[r != r] = 0
        refs.append(r)
    qdist.scrub_nan_grads(params)
    for p, r in zip(params[:-1], refs):
        assert torch.equal(p.grad, r)


@pytest.mark.parametrize("scaling", ["batch", "whole"])
def test_fused_elbo_matches_the_reference_expressions(scaling, golden):
    """losses.ClassificationLoss on CUDA tensors (one launch: qbn_elbo_cls) == src/losses.py:14-29 evaluated with torch on the same
    tensors — loss, both terms, d/d output and d/d kl — and == the reference-generated fixture."""
    import torch.nn.functional as F
    from qbn_b200 import losses, zoo
    g = torch.Generator().manual_seed(12)
    B, K = 37, 10
    out = torch.softmax(torch.randn(B, K, generator=g), -1).cuda().requires_grad_(True)
    tgt = torch.randint(0, K, (B,), generator=g).cuda()
    kl = torch.tensor(1234.5, device="cuda", requires_grad=True)
    args = zoo.Args(loss_multiplier=0.7)
    crit = losses.LOSS_FACTORY["classification"](args, scaling)
    gamma, n_batches, n_points = 0.3, 176, 45000
    loss, data, klt = crit(out, tgt, kl, gamma, n_batches, n_points)
    (loss + 0.5 * data + 0.25 * klt).backward()
    o2, k2 = out.detach().clone().requires_grad_(True), kl.detach().clone().requires_grad_(True)
    if scaling == "whole":
        ce = n_points * F.nll_loss(torch.log(o2 + 1e-8), tgt) * args.loss_multiplier
        kk = k2 / n_batches
    else:
        ce = F.nll_loss(torch.log(o2 + 1e-8), tgt)
        kk = k2 / (B * n_batches)
    ref = ce + gamma * kk
    (ref + 0.5 * ce + 0.25 * kk).backward()
    for a, b in ((loss, ref), (data, ce), (klt, kk), (out.grad, o2.grad), (kl.grad, k2.grad)):
        np.testing.assert_allclose(a.detach().cpu().numpy(), b.detach().cpu().numpy(), rtol=2e-6, atol=1e-7)
    # the fixture generated by the unmodified reference (oracle/make_golden.py:gen_losses): loss_multiplier 0.5, kl 123.4, gamma .01
    f = golden("losses")
    o = torch.as_tensor(f["out"]).cuda().requires_grad_(True)
    vals = losses.LOSS_FACTORY["classification"](zoo.Args(loss_multiplier=0.5), scaling)(o, torch.as_tensor(f["target"]).cuda(), torch.tensor(123.4).cuda(),
                                                                                      0.01, 176, 45000)
    vals[0].backward()
    np.testing.assert_allclose([float(v) for v in vals], f["cls_%s" % scaling], rtol=2e-6)
    np.testing.assert_allclose(o.grad.cpu().numpy(), f["cls_%s_dout" % scaling], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("geom", [(3, 24, 12, 20, 24, 3, 1, 1), (2, 16, 10, 6, 32, 3, 2, 1), (1, 8, 5, 9, 8, 3, 1, 1), (2, 32, 8, 14, 16, 1, 2, 0),
                                  (1, 24, 32, 32, 24, 3, 1, 1)])
def test_lrt_p4_non_square_maps_and_batch_one(geom):
    """The planar LRT path on H != W maps and B = 1, through ops.LRTFunction, against the oracle (forward, dx, d_mu, d_rho)."""
    from qbn_b200 import ops
    B, C, H, W, N, k, stride, pad = geom
    g = torch.Generator().manual_seed(31 + H * W)
    x = torch.randn(B, C, H, W, generator=g)
    mu = torch.randn(N, C, k, k, generator=g) / (C * k * k) ** 0.5
    rho = torch.empty(N, C, k, k).uniform_(-5, -2, generator=g)
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    eps = torch.randn(B, N, Ho, Wo, generator=g)
    go = torch.randn(B, N, Ho, Wo, generator=g)
    xx, m, r = x.cuda().requires_grad_(True), mu.cuda().requires_grad_(True), rho.cuda().requires_grad_(True)
    calls = []
    real = ops._lib.call
    ops._lib.call = lambda name, *a: (calls.append(name), real(name, *a))[1]
    try:
        out = ops.LRTFunction.apply(xx, m, r, None, (stride, stride), (pad, pad), (1, 1), eps.cuda(), (0, 0, 0), ops.QBN_MATH_TF32, False, None)
        (out * go.cuda()).sum().backward()
    finally:
        ops._lib.call = real
    assert "qbn_lrt_conv_p4_fwd" in calls and "qbn_lrt_wgrad_p4" in calls
    yo, so = O.lrt_conv_fwd(x, mu, rho, None, eps, stride, pad)
    dx_o, dmu_o, drho_o, _ = O.lrt_conv_bwd(x, mu, rho, eps, so, go, stride, pad)
    close(out, yo)
    close(xx.grad, dx_o)
    close(m.grad, dmu_o)
    close(r.grad, drho_o)
