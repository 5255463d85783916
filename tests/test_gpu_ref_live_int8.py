"""GPU, live against the unmodified reference (oracle/_ref on the GPU box, /root/reference in the build container; skipped where
neither exists): the reference's stock-quantised families of config C5 — the MC-Dropout ResNet-18 and the SGHMC ensemble —
built and converted by the reference's own lifecycle on the CPU (FBGEMM), then re-housed on the CUDA kernels
(quant_utils.to_device_int8) and evaluated (i) by the reference's own model code on QTensors, (ii) by the engines
(Int8EnsembleEngine on the planar kind::i8 kernel, Int8MCEngine with all samples of a chunk per launch).  Integer arithmetic:
the class probabilities must agree to fp32 softmax rounding (a single differing int8 logit would move them by ~1e-2)."""
import numpy as np
import pytest
import torch

import _ref_models as R

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not R.available(), reason="reference sources not present (oracle/_ref)")]


def _build():
    import __graft_entry__ as ge
    ge.build()


def test_sghmc_ensemble_int8_matches_the_reference_on_fbgemm():
    _build()
    from qbn_b200 import dist as qdist, quant_utils as qu
    from qbn_b200.mc_int8 import Int8EnsembleEngine, Int8PlanarEngine
    torch.set_num_threads(1)
    n = 3
    net, args = R.sgld_ensemble(n)
    x = torch.randn(8, 3, 32, 32, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        want = torch.stack([net(x) for _ in range(n)])                  # experiments/utils.py:344-347 over Network.forward
    mine = qu.to_device_int8(R.clone(net), "cuda")
    with torch.no_grad():
        got = torch.stack([mine(x.cuda()) for _ in range(n)])            # the reference's own forward code on the CUDA modules
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-6)
    eng = Int8EnsembleEngine(mine)
    assert all(isinstance(e, Int8PlanarEngine) for e in eng.engines)     # every member compiled onto the planar kernel
    p = eng.predict(x.cuda())
    np.testing.assert_allclose(p.cpu().numpy(), want.mean(0).numpy(), rtol=1e-5, atol=1e-6)
    assert torch.equal(p, eng.predict(x.cuda()))                         # graph replay
    # member sharding = sample sharding (SURVEY 8e iii): any split of the member range sums to the same probabilities
    parts = eng.predict_sum(x.cuda(), 2, sample0=0) + eng.predict_sum(x.cuda(), 1, sample0=2)
    np.testing.assert_allclose((parts / n).cpu().numpy(), p.cpu().numpy(), rtol=0, atol=1e-6)
    assert qdist.ShardedMCPredictor(eng).predict(x.cuda(), n).shape == (8, 10)


def test_mc_dropout_resnet_int8_matches_the_reference_on_fbgemm():
    _build()
    from qbn_b200 import noise, quant_utils as qu
    from qbn_b200.mc_int8 import Int8MCEngine
    torch.set_num_threads(1)
    net, args = R.mc_dropout_resnet()
    x = torch.randn(8, 3, 32, 32, generator=torch.Generator().manual_seed(4))
    sites = R.dropout_sites(net)
    shapes = []
    hooks = [m.register_forward_hook(lambda mod, i, o: shapes.append(tuple(i[0].shape[:2]))) for m in sites]
    torch.manual_seed(78)
    with torch.no_grad():
        want = net(x)
    for h in hooks:
        h.remove()
    torch.manual_seed(78)
    masks = [torch.FloatTensor(*s).bernoulli_(1. - sites[0].p).cuda() for s in shapes]      # dropout.py:21-30, same generator stream
    mine = qu.to_device_int8(R.clone(net), "cuda")
    with torch.no_grad(), noise.inject(masks):
        got = mine(x.cuda())
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-6)
    # product path: Philox masks, all samples of a chunk per launch == the per-sample loop with the same streams
    noise.manual_seed(5)
    S = 5
    with torch.no_grad():
        loop = []
        for s in range(S):
            with noise.sample_index(s):
                loop.append(mine(x.cuda()))
    assert float((loop[0] - loop[1]).abs().max()) > 0
    eng = Int8MCEngine(mine, chunk=3)
    got_sum = eng.predict_sum(x.cuda(), S)
    np.testing.assert_allclose(got_sum.cpu().numpy(), torch.stack(loop).sum(0).cpu().numpy(), rtol=0, atol=2e-6)
    rate = float(torch.stack([(m == 0).float().mean() for m in masks]).mean())
    assert 0.05 < rate < 0.25                                            # p = 0.15
    # the planar kind::i8 engine takes the MC-Dropout network too (one elementwise launch per site: mask multiply + the block's
    # residual add): same Philox draws, same integers as the module-driven engine -> identical probability sums
    from qbn_b200.mc_int8 import Int8PlanarEngine, make_int8_engine
    pe = make_int8_engine(mine, chunk=3)
    assert isinstance(pe, Int8PlanarEngine) and sum(st.dropout is not None for st in pe.steps) == len(sites)
    got_planar = pe.predict_sum(x.cuda(), S)
    np.testing.assert_allclose(got_planar.cpu().numpy(), got_sum.cpu().numpy(), rtol=0, atol=2e-6)
    assert torch.equal(pe.predict_sum(x.cuda(), S), got_planar)          # graph replay
    one = Int8PlanarEngine(mine, chunk=1, use_graph=False).predict_sum(x.cuda(), S)      # chunking does not change the draws
    np.testing.assert_allclose(one.cpu().numpy(), got_planar.cpu().numpy(), rtol=0, atol=2e-6)
    # unit-window sharding (dist.shard_units) over 3 emulated ranks: the fixed-weight first conv on the shared input runs for the
    # whole batch (a launch over one sample is outside the window of a 2-sample chunk), the rest only on the rank's rows
    from qbn_b200.dist import shard_units
    tot = torch.zeros_like(got_planar)
    for r in range(3):
        s0, n, first, end = shard_units(S, 8, r, 3)
        tot += pe.predict_sum(x.cuda(), n, sample0=s0, window=(first, end))
    np.testing.assert_allclose(tot.cpu().numpy(), got_planar.cpu().numpy(), rtol=0, atol=2e-6)


def test_sghmc_optimiser_step_matches_the_reference():
    """The fused SGHMC update (qbn_sghmc_step behind the drop-in SGLD optimiser) against the reference's SGLD class
    (utils_sgld.py:30-92) on the CPU, with its Gaussian draws replayed (torch.normal(0, std) == normal_() * std on the same
    generator stream): parameters and every state tensor after a burn-in step with momentum resampling, a burn-in step and a
    sampling step."""
    _build()
    from oracle import ref_harness
    ref_harness.import_reference()
    from src.models.stochastic.sgld.utils_sgld import SGLD as RefSGLD
    from qbn_b200 import noise
    from qbn_b200.stochastic.sgld.utils_sgld import SGLD
    g = torch.Generator().manual_seed(9)
    shapes = [(64, 33), (17,), (8, 4, 3, 3)]
    p_ref = [torch.nn.Parameter(torch.randn(s, generator=g) * 0.3) for s in shapes]
    p_gpu = [torch.nn.Parameter(p.detach().clone().cuda()) for p in p_ref]
    o_ref, o_gpu = RefSGLD(p_ref, lr=1e-2, base_C=0.05, gauss_sig=0.1), SGLD(p_gpu, lr=1e-2, base_C=0.05, gauss_sig=0.1)
    plan = [dict(burn_in=True, resample_momentum=True), dict(burn_in=True, resample_momentum=False), dict(burn_in=False, resample_momentum=False)]
    for k, kw in enumerate(plan):
        grads = [torch.randn(s, generator=g) for s in shapes]
        for p, q, gr in zip(p_ref, p_gpu, grads):
            p.grad, q.grad = gr.clone(), gr.clone().cuda()
        torch.manual_seed(100 + k)
        o_ref.step(**kw)
        torch.manual_seed(100 + k)
        zs = []
        for s in shapes:                                   # per parameter: [momentum draw,] noise draw — the reference's order
            if kw["resample_momentum"]:
                zs.append(torch.empty(s).normal_().cuda())
            zs.append(torch.empty(s).normal_().cuda())
        with noise.inject(zs):
            o_gpu.step(**kw)
        for p, q in zip(p_ref, p_gpu):
            np.testing.assert_allclose(q.detach().cpu().numpy(), p.detach().numpy(), rtol=2e-6, atol=1e-7)
            np.testing.assert_allclose(q.grad.cpu().numpy(), p.grad.numpy(), rtol=2e-6, atol=1e-7)      # weight decay folded in place
            for name in ("tau", "g", "V_hat", "v_momentum"):
                # tau += -tau g^2 / (V + eps) + 1 cancels to ~1e-6 per step: an ulp of g or V_hat is amplified, hence 2e-5 on tau
                np.testing.assert_allclose(o_gpu.state[q][name].cpu().numpy(), o_ref.state[p][name].numpy(), rtol=2e-5 if name == "tau" else 2e-6,
                                           atol=1e-7, err_msg=name)
    # product path: Philox noise — finite, and different from step to step
    for q in p_gpu:
        q.grad = torch.randn_like(q)
    before = [q.detach().clone() for q in p_gpu]
    o_gpu.step(burn_in=False, resample_momentum=True)
    assert all(torch.isfinite(q).all() and not torch.equal(q, b) for q, b in zip(p_gpu, before))
