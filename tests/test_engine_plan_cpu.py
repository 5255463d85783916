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


def test_p4map_layout_roundtrip_cpu():
    """The planar-C4 layout with shared zero borders (pure torch, CPU): round trips, zero rows on top / zero columns on the
    left of every map, zero tail, and the phase-split storage of a map for a stride-2 consumer."""
    from qbn_b200.ops import P4Map
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 8, 6, 4, generator=g)
    m = P4Map.from_nchw(x, (1, 1))
    assert (m.Hp, m.Wp, m.phases) == (7, 5, 1) and m.plane_rows == 3 * 7 * 5 + 5 + 1 and m.buf.shape == (2, m.plane_rows, 4)
    assert torch.equal(m.to_nchw(), x)
    full = m.to_nchw(keep_border=True)
    assert float(full[:, :, 0, :].abs().max()) == 0.0 and float(full[:, :, :, 0].abs().max()) == 0.0 and float(m.tail().abs().max()) == 0.0
    # flat-index property the kernel relies on: pixel (b, h, w) sits at row b*Hp*Wp + (h+1)*Wp + (w+1) of every chunk plane,
    # and its (dr, ds) neighbour at + dr*Wp + ds — zero whenever it falls outside the map
    rows = m.buf.permute(1, 0, 2).reshape(m.plane_rows, 8)
    b, h, w = 1, 5, 3                       # bottom-right pixel of image 1
    q = b * 35 + (h + 1) * 5 + (w + 1)
    assert torch.equal(rows[q], x[b, :, h, w])
    assert float(rows[q + 1].abs().max()) == 0.0 and float(rows[q + 5].abs().max()) == 0.0 and float(rows[q + 6].abs().max()) == 0.0
    assert torch.equal(rows[q - 6], x[b, :, h - 1, w - 1])
    q_last = 2 * 35 + 6 * 5 + 4              # bottom-right pixel of the LAST image: its bottom/right neighbours are the tail
    assert q_last + 6 < m.plane_rows and float(rows[q_last + 1:].abs().max()) == 0.0
    # phase split: phase (a, b) holds pixels (2i+a, 2j+b) at (i+1, j+1) of a (H/2+1) x (W/2+1) map
    ps = P4Map.from_nchw(x, None, phase_split=True)
    assert (ps.Hp, ps.Wp, ps.phases) == (4, 3, 4) and torch.equal(ps.to_nchw(), x)
    prow = ps.buf.permute(1, 0, 2).reshape(ps.plane_rows, 8)
    phase, i, j = 1 * 2 + 0, 2, 1            # pixel (2*2+1, 2*1+0) = (5, 2)
    assert torch.equal(prow[phase * 3 * 12 + 0 * 12 + (i + 1) * 3 + (j + 1)], x[0, :, 5, 2])
