"""CPU: the MC engine's compile step and planar-layout planner (pure host logic; no kernels run)."""
import torch

import oracle.qbn_oracle as O


def _engine(net, **kw):
    from qbn_b200 import mc
    return mc, mc.MCEngine(net.eval(), math_mode="tf32", **kw)


def test_resnet_bbb_plan():
    from qbn_b200 import zoo
    mc, eng = _engine(zoo.resnet_from_params(O.ResNetBBBParams(seed=1)))
    layout, on = eng._plan_p4()
    convs = [st for st in eng.steps if isinstance(st, mc._ConvStep)]
    assert len(convs) == 21 and eng.n_noise == 21              # one noise draw per Bayesian layer, in the reference's order
    assert sum(1 for st in convs if id(st) in on) == 19          # every conv but the first (3 channels) and the classifier
    assert eng._p4_first == id(convs[0])                         # first layer: sample-stacked planar launch
    # the three downsampling shortcuts are accumulated inside the blocks' second stem conv
    assert len(eng._p4_fused) == 3 and len(eng._p4_skip) == 3
    for st_id, sc in eng._p4_fused.items():
        st = next(s for s in convs if id(s) == st_id)
        assert tuple(sc.mod.kernel_size) == (1, 1) and tuple(sc.mod.stride) == (2, 2) and st.residual == sc.dst
    # registers feeding a stride-2 conv are phase-split, the rest planar with a 1-pixel border
    kinds = [layout[st.dst][0] for st in convs if st.dst in layout]
    assert kinds.count("p4s") == 3 and all(k in ("p4", "p4s") for k in kinds)
    # shortcut first in execution order, reference order in the noise indices (stem.0, stem.3, shortcut.0)
    blk = [st for st in convs if not st.is_linear and st.mod.in_channels == 24 and st.mod.out_channels == 48]
    assert [tuple(st.mod.kernel_size) for st in blk] == [(1, 1), (3, 3)] and blk[0].ref_idx > blk[1].ref_idx


def test_resnet_mc_dropout_plan():
    from qbn_b200 import zoo
    P = O.ResNetBBBParams(seed=1)
    mc, eng = _engine(zoo.resnet_mc_from_params(P, 0.15, state_dict=O.resnet_mc_state_dict(P)))
    convs = [st for st in eng.steps if isinstance(st, mc._ConvStep)]
    assert all(st.det for st in convs)
    sites = [st for st in convs if st.dropout is not None]
    assert len(sites) == 20 and eng.n_noise == 20                # one mask draw per dropout site
    assert sorted(st.dropout[2] for st in sites) == list(range(20))
    first = convs[0]
    assert first.relu_pre and not first.relu                     # conv-BN-ReLU-dropout: the ReLU precedes the mask
    last_of_block = [st for st in convs if st.residual is not None]
    assert all(st.relu and not st.relu_pre and st.dropout is not None for st in last_of_block)   # conv-BN-dropout-add-ReLU
    layout, on = eng._plan_p4()
    assert sum(1 for st in convs if id(st) in on) == 19 and not eng._p4_fused      # dropout on the shortcut: no fusion


def test_lenet_plans_stay_on_the_gather_path():
    from qbn_b200 import zoo
    P = O.LeNetBBBParams(seed=3)
    for net in (zoo.lenet_from_params(P), zoo.lenet_mc_from_params(P, 0.2)):
        mc, eng = _engine(net)
        layout, on = eng._plan_p4()
        assert not on and not layout and eng._p4_first is None   # 1 / 20 input channels: not planar-eligible


def test_fp32_mode_never_plans_planar():
    from qbn_b200 import mc, zoo
    eng = mc.MCEngine(zoo.resnet_from_params(O.ResNetBBBParams(seed=1)).eval(), math_mode="fp32")
    layout, on = eng._plan_p4()
    assert not on and not layout


def test_balanced_chunks():
    # 13 samples with chunk 10 -> 7 + 6 (sharding 100 samples over 8 GPUs)
    n_chunks = (13 + 9) // 10
    sizes = [13 // n_chunks + (1 if i < 13 % n_chunks else 0) for i in range(n_chunks)]
    assert sizes == [7, 6]
