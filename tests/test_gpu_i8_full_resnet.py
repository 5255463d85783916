"""Config C5 at FULL size against the reference's own FBGEMM run (tests/golden/resnet_int8_full*.{npz,pt}, written by
oracle/make_golden.py:gen_full_resnet_int8 from the UNMODIFIED reference): the narrow ResNet-18 (24/48/96/192) quantised
A7/W8, loaded from the checkpoint the reference writes, B=4, two forwards with the reference's noise replayed.

  * module path (one forward per sample, NHWC kernels): the SHA-1 of EVERY int8 layer's integer output, every BasicBlock
    output, the pooled map, the logits and the class probabilities equal the reference's;
  * planar engine (qbn_i8_conv_p16_fwd, both samples in one launch per layer): every conv output it materialises, every
    BasicBlock output (residual add + ReLU fused in the epilogue), the pooled map, the logits and p-bar equal the reference's.
Integer maps are compared bit for bit."""
import hashlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def replay_eps(net, g, fi):
    """The eps tensors of forward `fi`, redrawn from torch's CPU generator exactly like the fixture script did."""
    q_names = [str(n) for n in g["q_names"]]
    mods = dict(net.named_modules())
    torch.manual_seed(int(g["seeds"][fi]))
    eps, h = [], hashlib.sha1()
    for n in q_names:
        e = torch.empty(tuple(mods[n].std.shape)).normal_()
        h.update(np.ascontiguousarray(e.numpy()).tobytes())
        eps.append(e)
    assert h.hexdigest() == str(g["f%d.eps_sha1" % fi]), "torch's CPU generator no longer reproduces the fixture's noise stream"
    return eps


def full_int8_resnet(golden_dir, device="cuda"):
    import __graft_entry__ as ge
    ge.build()
    from qbn_b200 import quant_utils as qu, zoo
    args = zoo.Args(sigma_prior=0.05, model="conv_resnet_bbb", q=True, at=True, activation_precision=7, weight_precision=8)
    net = zoo.ConvNetwork_ResNet([1, 3, 32, 32], 10, True, args)
    net.train()
    qu.prepare_model(net, args)
    qu.convert(net.to(device))                             # un-calibrated skeleton; every number comes from the reference's file
    qu.load_model(net, str(golden_dir / "resnet_int8_full_weights.pt"))
    return net.eval()


def test_full_resnet_module_path_reproduces_every_reference_integer_map(golden, golden_dir):
    from qbn_b200 import noise
    g = golden("resnet_int8_full")
    net = full_int8_resnet(golden_dir)
    assert set(net.state_dict().keys()) == set(torch.load(golden_dir / "resnet_int8_full_weights.pt", map_location="cpu").keys())
    x = torch.as_tensor(g["x"]).cuda()
    mods = dict(net.named_modules())
    order = [str(n) for n in g["order"]]
    for fi in (0, 1):
        seen = {}
        # (the model applies quantised pooling itself — zoo._apply — so the pooled map is caught as the Flatten's input)
        hooks = [mods[n].register_forward_hook(lambda m, i, o, n=n: seen.__setitem__(n, o)) for n in order if n != "layers.7"]
        hooks.append(mods["layers.8"].register_forward_pre_hook(lambda m, i: seen.__setitem__("layers.7", i[0])))
        with torch.no_grad(), noise.inject([e.cuda() for e in replay_eps(net, g, fi)]):
            y = net(x)
        for h in hooks:
            h.remove()
        for n in order:
            out = seen[n]
            ints = out.q.cpu().numpy()                     # logical NCHW, like int_repr() of the reference's tensor
            want_qp = g["f%d.%s.qp" % (fi, n)]
            assert out.zero_point == int(want_qp[1]) and abs(out.scale - float(want_qp[0])) < 1e-12, n
            key = "f%d.%s.y_q" % (fi, n)
            if key in g.files:
                assert np.array_equal(ints.reshape(g[key].shape), g[key]), n
            assert _sha(ints.astype(np.uint8)) == str(g["f%d.%s.sha1" % (fi, n)]), "integer map of %s differs from FBGEMM's" % n
        np.testing.assert_allclose(y.cpu().numpy(), g["f%d.y" % fi], rtol=1e-5, atol=1e-7)


def test_full_resnet_planar_engine_reproduces_the_reference(golden, golden_dir):
    from qbn_b200.mc_int8 import Int8PlanarEngine
    g = golden("resnet_int8_full")
    net = full_int8_resnet(golden_dir)
    eng = Int8PlanarEngine(net, chunk=2, use_graph=False)
    assert eng.n_noise == len(g["q_names"]) == 21
    x = torch.as_tensor(g["x"]).cuda()
    injected = [[e.cuda() for e in replay_eps(net, g, fi)] for fi in (0, 1)]
    eng.trace = {}
    psum = eng.predict_sum(x, 2, injected=injected)
    trace, eng.trace = eng.trace, None
    B = x.shape[0]
    order = [str(n) for n in g["order"]]
    checked = 0
    for st in eng.steps:
        ints, scale, zp = trace[st.name]
        for fi in (0, 1):
            mine = ints[fi * B:(fi + 1) * B].cpu().numpy()
            if st.residual is None:                        # a conv whose own output exists in the reference too
                want_qp = g["f%d.%s.qp" % (fi, st.name)]
                assert zp == int(want_qp[1]) and abs(scale - float(want_qp[0])) < 1e-12, st.name
                # the epilogue already applied the clamp_activation that follows every module (models_bbb.py:231-238)
                assert _sha(mine.astype(np.uint8)) == str(g["f%d.%s.sha1_clamped" % (fi, st.name)]), "conv output of %s differs from FBGEMM's" % st.name
            else:                                          # second stem conv: its epilogue wrote the BasicBlock's output
                blk = st.name.rsplit(".stem.", 1)[0]
                assert blk in order
                want = g["f%d.%s.y_q" % (fi, blk)]
                assert np.array_equal(mine, want), "output of block %s differs from the reference's (%d of %d)" % (blk, int((mine != want).sum()), want.size)
            checked += 1
    assert checked == 2 * 20
    pooled = trace["pool"][0].cpu().numpy().reshape(2, B, -1)
    logits = trace["head"][0].cpu().numpy().reshape(2, B, -1)
    for fi in (0, 1):
        assert np.array_equal(pooled[fi], g["f%d.layers.7.y_q" % fi].reshape(B, -1)), "pooled map"
        assert np.array_equal(logits[fi], g["f%d.layers.9.y_q" % fi].reshape(B, -1)), "int8 logits"
    np.testing.assert_allclose(psum.cpu().numpy(), g["f0.y"] + g["f1.y"], rtol=1e-5, atol=1e-6)


def test_planar_engine_equals_the_module_driven_engine_with_philox_noise(golden_dir):
    """Same Philox streams, same integers: the planar engine (graph replay, any chunking, any first sample index) returns the
    probabilities of the module-driven engine on the exact CUDA-core kernels."""
    from qbn_b200 import noise
    from qbn_b200.mc_int8 import Int8MCEngine, Int8PlanarEngine, make_int8_engine
    net = full_int8_resnet(golden_dir)
    x = torch.randn(8, 3, 32, 32, generator=torch.Generator().manual_seed(5)).cuda()
    noise.manual_seed(99)
    ref = Int8MCEngine(net, chunk=3, tensor_cores=False).predict_sum(x, 5, sample0=2)
    fast = Int8PlanarEngine(net, chunk=3)
    assert isinstance(make_int8_engine(net), Int8PlanarEngine)
    got = fast.predict_sum(x, 5, sample0=2)
    np.testing.assert_allclose(got.cpu().numpy(), ref.cpu().numpy(), rtol=0, atol=2e-6)       # same ints; fp32 summation order only
    again = fast.predict_sum(x, 5, sample0=2)                                                # graph replay
    assert torch.equal(got, again)
    other = Int8PlanarEngine(net, chunk=5, use_graph=False).predict_sum(x, 5, sample0=2)
    np.testing.assert_allclose(other.cpu().numpy(), ref.cpu().numpy(), rtol=0, atol=2e-6)
    p = fast.predict(x, 4)
    np.testing.assert_allclose(p.sum(-1).cpu().numpy(), np.ones(8), atol=1e-5)
    # unit-window sharding (dist.shard_units): 5 samples x 8 images over 4 and 16 emulated ranks add up to the unsharded sum
    from qbn_b200.dist import shard_units
    full = fast.predict_sum(x, 5, sample0=0)
    for world in (4, 16):
        tot = torch.zeros_like(full)
        for r in range(world):
            s0, n, first, end = shard_units(5, 8, r, world)
            if n:
                tot += fast.predict_sum(x, n, sample0=s0, window=(first, end))
        np.testing.assert_allclose(tot.cpu().numpy(), full.cpu().numpy(), rtol=0, atol=2e-6)
