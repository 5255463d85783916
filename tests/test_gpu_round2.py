"""GPU (pytest -m gpu): round-2 boundary behaviour — gradients through MC-Dropout and through the eval-mode (weight-sampling)
forward, the reference-signature KL, FakeQuantize <-> torch's FakeQuantize checkpoints, the device-side draw offset (one CUDA
graph, fresh noise per batch), the input-width gate of the tcgen05 int8 kernel, the metric containers' reference surface,
sharded regression, and a full-size (B=256) spot check of the engine against the oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _lib():
    import __graft_entry__ as g
    g.build()


def test_gradients_reach_a_conv_in_front_of_mc_dropout():
    """dropout.py:35-39 is differentiable (x * mask * 1/(1-p)); the reference trains MC-Dropout nets through it."""
    from qbn_b200 import noise
    from qbn_b200.stochastic.mcdropout.dropout import BernoulliDropout
    g = torch.Generator().manual_seed(1)
    conv = torch.nn.Conv2d(3, 8, 3, padding=1).cuda()
    drop = BernoulliDropout(0.25).cuda()
    x = torch.randn(4, 3, 6, 6, generator=g).cuda()
    mask = torch.empty(4, 8).bernoulli_(0.75, generator=g).cuda()
    with noise.inject([mask]):
        y = drop(conv(x))
    y.square().sum().backward()
    ref_conv = torch.nn.Conv2d(3, 8, 3, padding=1).cuda()
    ref_conv.load_state_dict(conv.state_dict())
    z = ref_conv(x) * mask.view(4, 8, 1, 1) * (1.0 / (1.0 - 0.25))
    z.square().sum().backward()
    np.testing.assert_allclose(y.detach().cpu().numpy(), z.detach().cpu().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(conv.weight.grad.cpu().numpy(), ref_conv.weight.grad.cpu().numpy(), rtol=1e-4, atol=1e-5)
    assert float(conv.weight.grad.abs().sum()) > 0
    # p follows the state-dict (a checkpoint may change it after construction)
    drop.load_state_dict({"p": torch.ones(1) * 0.5, "multiplier": torch.ones(1) * 2.0}, strict=False)
    assert drop._prob() == 0.5


def test_eval_mode_forward_is_differentiable_like_the_reference():
    """linear.py:42-50 / conv.py:33-39 are plain autograd in the reference: dx, dmu, drho of y = conv(x, mu + softplus(rho) eps)."""
    from qbn_b200 import config, noise
    from qbn_b200.stochastic.bbb.conv import Conv2d
    config.set_math_mode("fp32")
    try:
        g = torch.Generator().manual_seed(2)
        m = Conv2d(4, 6, (3, 3), stride=1, padding=1, bias=False, sigma_prior=0.1).cuda().eval()
        with torch.no_grad():
            m.weight.copy_(torch.randn(m.weight.shape, generator=g) * 0.2)
            m.std.copy_(torch.empty(m.std.shape).uniform_(-3, -1, generator=g))
        x = torch.randn(2, 4, 5, 5, generator=g).cuda().requires_grad_(True)
        eps = torch.randn(m.weight.shape, generator=g).cuda()
        with noise.inject([eps]):
            y = m(x)
        y.square().sum().backward()
        mu, rho, xr = m.weight.detach().clone().requires_grad_(True), m.std.detach().clone().requires_grad_(True), x.detach().clone().requires_grad_(True)
        z = F.conv2d(xr, mu + F.softplus(rho) * eps, None, 1, 1)
        z.square().sum().backward()
        np.testing.assert_allclose(y.detach().cpu().numpy(), z.detach().cpu().numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(x.grad.cpu().numpy(), xr.grad.cpu().numpy(), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(m.weight.grad.cpu().numpy(), mu.grad.cpu().numpy(), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(m.std.grad.cpu().numpy(), rho.grad.cpu().numpy(), rtol=1e-4, atol=1e-5)
    finally:
        config.set_math_mode("tf32")


def test_kl_divergence_with_the_reference_argument_list():
    """utils_bbb.py:3-5 called the way linear.py:24-28 calls it: (mu, softplus(rho), zeros_like, ones_like * std_prior)."""
    from qbn_b200.stochastic.bbb.utils_bbb import kl_divergence, kl_divergence_from_rho
    g = torch.Generator().manual_seed(3)
    mu = (torch.randn(50, 30, generator=g) * 0.1).cuda().requires_grad_(True)
    rho = torch.empty(50, 30).uniform_(-4, -1, generator=g).cuda().requires_grad_(True)
    prior = torch.ones(1).cuda() * 0.3
    kl = kl_divergence(mu, F.softplus(rho), torch.zeros_like(mu), torch.ones_like(rho) * prior)
    kl.backward()
    mu2, rho2 = mu.detach().clone().requires_grad_(True), rho.detach().clone().requires_grad_(True)
    sg = F.softplus(rho2)
    ref = 0.5 * (2 * torch.log(prior / sg) - 1 + (sg / prior).pow(2) + ((0 - mu2) / prior).pow(2)).sum()
    ref.backward()
    np.testing.assert_allclose(float(kl), float(ref), rtol=1e-5)
    np.testing.assert_allclose(mu.grad.cpu().numpy(), mu2.grad.cpu().numpy(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(rho.grad.cpu().numpy(), rho2.grad.cpu().numpy(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(float(kl_divergence_from_rho(mu.detach(), rho.detach(), prior)), float(ref), rtol=1e-5)
    with pytest.raises(NotImplementedError):
        kl_divergence(mu, F.softplus(rho), torch.randn_like(mu), prior)          # a non-constant prior mean is outside the path


def test_fake_quantize_checkpoints_are_torchs():
    """A QAT checkpoint written by torch's FakeQuantize(MovingAverageMinMaxObserver) (what the reference saves) restores the
    trained EMA range here, and the other way round; torch's enable/disable helpers reach the drop-in."""
    import torch.ao.quantization as taq
    from qbn_b200.quant_utils import FakeQuantize
    ref = taq.FakeQuantize(observer=taq.MovingAverageMinMaxObserver, quant_min=0, quant_max=127, dtype=torch.quint8, qscheme=torch.per_tensor_affine)
    g = torch.Generator().manual_seed(4)
    xs = [torch.randn(64, 32, generator=g) * (1 + i) for i in range(3)]
    for x in xs:
        ref(x)
    mine = FakeQuantize(quant_min=0, quant_max=127, dtype=torch.quint8).cuda()
    assert set(mine.state_dict().keys()) == set(ref.state_dict().keys())
    mine.load_state_dict(ref.state_dict())
    assert float(mine.activation_post_process.min_val) == float(ref.activation_post_process.min_val)
    x = torch.randn(64, 32, generator=g) * 2
    y_ref, y = ref(x), mine(x.cuda())                       # both take one more EMA step from the SAME trained state
    np.testing.assert_allclose(y.cpu().numpy(), y_ref.numpy(), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(float(mine.scale), float(ref.scale), rtol=1e-6)
    back = taq.FakeQuantize(observer=taq.MovingAverageMinMaxObserver, quant_min=0, quant_max=127, dtype=torch.quint8, qscheme=torch.per_tensor_affine)
    back.load_state_dict({k: v.cpu() for k, v in mine.state_dict().items()})
    np.testing.assert_allclose(float(back.activation_post_process.max_val), float(ref.activation_post_process.max_val), rtol=1e-6)
    holder = torch.nn.Sequential(mine)
    holder.apply(taq.disable_observer)
    assert mine._observer_on is False and int(mine.observer_enabled) == 0
    holder.apply(taq.disable_fake_quant)
    lo = float(mine.activation_post_process.min_val)
    assert torch.equal(mine(x.cuda() * 100), x.cuda() * 100) and float(mine.activation_post_process.min_val) == lo
    holder.apply(taq.enable_observer)
    mine(x.cuda() * 100)                                     # fake-quant off, observer on: torch still observes
    assert float(mine.activation_post_process.min_val) < lo


def test_one_graph_serves_fresh_noise_per_batch():
    """mc.py: the captured graph reads the draw offset from a device scalar: batch k gets the sample indices k*S .. k*S+S-1
    without a re-capture, equal to an eager pass that starts at that sample index."""
    from qbn_b200 import mc, noise, synthetic, zoo
    model = zoo.resnet_from_params(synthetic.ResNetBBBParams(seed=3)).cuda().eval()
    x = torch.randn(4, 3, 32, 32, generator=torch.Generator().manual_seed(6)).cuda()
    noise.manual_seed(11)
    try:
        eng = mc.MCEngine(model, math_mode="tf32", chunk=4)
        S = 4
        p0 = eng.predict(x, S, draw_offset=0)
        p1 = eng.predict(x, S, draw_offset=S)
        p0_again = eng.predict(x, S, draw_offset=0)
        assert len(eng.__dict__["_graphs"]) == 1
        assert torch.equal(p0, p0_again) and not torch.equal(p0, p1)
        noise.set_draw_offset(0, x.device)
        eager = mc.MCEngine(model, math_mode="tf32", chunk=4, use_graph=False)
        np.testing.assert_allclose(p1.cpu().numpy(), eager.predict(x, S, sample0=S).cpu().numpy(), rtol=0, atol=1e-6)
        np.testing.assert_allclose(p0.cpu().numpy(), eager.predict(x, S, sample0=0).cpu().numpy(), rtol=0, atol=1e-6)
    finally:
        noise.set_draw_offset(0, x.device)


def test_unclamped_8bit_input_stays_off_the_7bit_tensor_core_kernel():
    """The tcgen05 kind::i8 kernel stages (x - z_x) as s8: a quint8 input that uses the full 8-bit range must take the exact
    CUDA-core kernel even when the layer's OUTPUT is clamped to 7 bits."""
    import oracle.qbn_oracle as O
    from qbn_b200 import ops
    rng = np.random.default_rng(8)
    B, C, H, N = 2, 16, 6, 8
    xq = rng.integers(0, 256, size=(B, C, H, H)).astype(np.uint8)
    w = rng.integers(-128, 128, size=(N, C, 3, 3)).astype(np.int8)
    want, _ = O.i8_conv(xq.astype(np.int32), 0.05, 3, w.astype(np.int32), 0.01, 0, None, 0.4, 10, 1, 1, 1, False, act_bits=7)
    x = torch.as_tensor(xq).cuda().contiguous(memory_format=torch.channels_last)
    wp = torch.as_tensor(w).cuda().permute(0, 2, 3, 1).contiguous().reshape(1, -1)
    d = ops.make_desc(B, H, H, C, N, 3, 3, 1, 1, 1)
    got = ops.i8_conv_forward(x, 0.05, 3, wp, 0.01, 0, d, None, 0.4, 10, False, act_bits=7, x_bits=8)
    assert np.array_equal(got.cpu().numpy().astype(np.int32), want)


def test_metric_containers_have_the_reference_surface(golden):
    """metric.error.compute().item() etc. (experiments/utils.py:372-375) and update(output=, target=, obj=, kl=, main_obj=)
    (trainer.py:128): values against the reference-generated golden sums."""
    from qbn_b200 import metrics
    g = golden("metrics")
    probs, target = torch.as_tensor(g["probs"]).cuda(), torch.as_tensor(g["target"]).cuda()
    m = metrics.ClassificationMetric(output_size=probs.shape[1])
    m.update(output=probs, target=target, obj=torch.tensor(2.0).cuda(), kl=torch.tensor(0.5).cuda(), main_obj=torch.tensor(1.5).cuda())
    m.update(output=probs, target=target, obj=torch.tensor(4.0).cuda(), kl=torch.tensor(1.5).cuda(), main_obj=torch.tensor(2.5).cuda())
    for name in ("error", "nll", "brier", "entropy", "ece"):
        np.testing.assert_allclose(getattr(m, name).compute().item(), float(g[name]), rtol=2e-5, atol=1e-6, err_msg=name)
    assert m.get_key_metric().item() == pytest.approx(float(g["error"]), abs=1e-6)
    packed = m.get_packed()
    assert packed["obj"] == pytest.approx(3.0) and packed["kl"] == pytest.approx(1.0) and packed["main_obj"] == pytest.approx(2.0)
    assert "Expected Calibration Error" in m.get_str()
    m.reset()
    assert m.count == 0 and m.obj.cnt == 0


def test_sharded_predictor_handles_regression_and_idle_ranks():
    from qbn_b200 import dist as qdist, mc, noise, zoo
    args = zoo.Args(sigma_prior=1.0, model="linear_bbb", task="regression")
    net = zoo.LinearNetwork([1], 1, False, args).cuda().eval()
    noise.manual_seed(5)
    eng = mc.MCEngine(net, math_mode="fp32", chunk=8)
    x = torch.linspace(-2, 2, 32).reshape(32, 1).cuda()
    mean, var = qdist.ShardedMCPredictor(eng).predict(x, 8)
    noise.manual_seed(5)
    m2, v2 = eng.predict(x, 8)
    np.testing.assert_allclose(mean.cpu().numpy().reshape(-1), m2.cpu().numpy().reshape(-1), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(var.cpu().numpy().reshape(-1), v2.cpu().numpy().reshape(-1), rtol=1e-3, atol=1e-5)


def test_full_batch_engine_spot_check_against_the_oracle():
    """BASELINE size in the batch dimension (B=256), two injected samples: p-bar of the planar TF32 engine against the oracle's
    CPU restatement of the reference forward with the same noise (TF32: 2e-3 absolute on probabilities, measured ~3e-4)."""
    import oracle.qbn_oracle as O
    from qbn_b200 import mc, zoo
    P = O.ResNetBBBParams(seed=21)
    x = torch.randn(256, 3, 32, 32, generator=torch.Generator().manual_seed(23))
    plan = O.resnet_noise_plan(P)
    noises = [O.replay_noise(900 + s, [p[1] for p in plan]) for s in range(2)]
    torch.set_num_threads(max(1, torch.get_num_threads()))
    ref = O.mc_mean_probs([O.resnet_bbb_eval_forward(P, x, lambda n, sh, d=dict(zip([p[0] for p in plan], nz)): d[n]) for nz in noises])
    eng = mc.MCEngine(zoo.resnet_from_params(P).cuda().eval(), math_mode="tf32")
    got = eng.predict(x.cuda(), 2, injected=[[t.cuda() for t in nz] for nz in noises]).cpu()
    err = (got - ref).abs()
    assert float(err.max()) < 2e-3, "max |p_gpu - p_oracle| = %g at B=256" % float(err.max())
    assert float(err.mean()) < 1e-4
