"""int8 on the planar zero-copy tcgen05 kernel (qbn_i8_conv_p16_fwd and its layout helpers) against the oracle's restatement
of FBGEMM's arithmetic (oracle/qbn_oracle.py: i8_conv = quantized::conv2d[_relu], qadd = quantized::add, i8_avgpool).
Integer work: every comparison is bit-exact."""
import numpy as np
import pytest
import torch

import oracle.qbn_oracle as O

pytestmark = pytest.mark.gpu


def _ops():
    import __graft_entry__ as ge
    ge.build()
    from qbn_b200 import ops
    return ops


def _rand_q(rng, shape, bits=7):
    return rng.integers(0, 1 << bits, size=shape, dtype=np.int64).astype(np.uint8)


def _to_map(ops, q_nchw, scale, zp, phase_split=False):
    """uint8 NCHW numpy -> P16Map on the GPU (phase-split maps are built with torch ops: the kernels only write them)."""
    t = torch.as_tensor(q_nchw).cuda().contiguous(memory_format=torch.channels_last)
    if not phase_split:
        return ops.P16Map.from_quint8(t, scale, zp, 7)
    n, C, H, W = q_nchw.shape
    m = ops.P16Map.empty(n, C, H // 2 + 1, W // 2 + 1, 4, "cuda", scale, zp, 7)
    xs = (t.permute(0, 2, 3, 1).to(torch.int16) - zp).to(torch.int8)            # [n, H, W, C] as q - z
    xs = torch.nn.functional.pad(xs, (0, m.C_pad - C))
    body = n * m.Hp * m.Wp
    k = 0
    for a in (0, 1):
        for b in (0, 1):
            ph = torch.nn.functional.pad(xs[:, a::2, b::2, :], (0, 0, 1, 0, 1, 0)).reshape(body, m.C_pad // 16, 16)
            m.buf[:, k * body:(k + 1) * body] = ph.permute(1, 0, 2)
            k += 1
    return m


CASES = [
    # C, N, H, k, stride, relu, residual, x_shared, out_split
    (24, 24, 8, 3, 1, True, False, False, False),
    (24, 24, 8, 3, 1, False, True, False, False),
    (3, 24, 8, 3, 1, True, False, True, False),          # first layer: one shared input, channels padded 3 -> 32
    (24, 24, 8, 3, 1, False, True, False, True),         # block output feeding a stride-2 block: phase-split store
    (24, 48, 8, 3, 2, True, False, False, False),        # stride-2 3x3 on a phase-split map
    (24, 48, 8, 1, 2, False, False, False, False),       # 1x1 stride-2 shortcut
    (48, 48, 4, 3, 1, False, True, False, False),
    (96, 96, 4, 3, 1, True, False, False, False),
    (192, 192, 4, 3, 1, False, True, False, False),      # streamed weight blocks
    (96, 192, 4, 3, 2, True, False, False, False),
]


@pytest.mark.parametrize("C,N,H,k,stride,relu,with_res,x_shared,out_split", CASES)
def test_i8_planar_conv_matches_fbgemm_arithmetic(C, N, H, k, stride, relu, with_res, x_shared, out_split):
    ops = _ops()
    rng = np.random.default_rng(1000 + C * 7 + N + H + k + stride + int(with_res) * 3 + int(out_split))
    B, n = 3, 2
    pad = (k - 1) // 2
    s_x, z_x = 0.043, int(rng.integers(0, 128))
    s_w, z_w = 0.0071, int(rng.integers(-20, 20))
    s_out, z_out = 0.37 if C > 48 else 0.11, int(rng.integers(0, 128))
    xq = _rand_q(rng, (B if x_shared else n * B, C, H, H))
    w = rng.integers(-128, 128, size=(n, N, C, k, k), dtype=np.int64).astype(np.int8)
    bias = rng.normal(0, 1.0, N).astype(np.float32)
    Ho = H // stride
    res_q = _rand_q(rng, (n * B, N, Ho, Ho)) if with_res else None
    s_res, z_res, s_add, z_add = 0.09, int(rng.integers(0, 128)), 0.21, int(rng.integers(0, 128))

    # ---- oracle, sample by sample
    want, want_acc = [], []
    for s in range(n):
        xs = xq if x_shared else xq[s * B:(s + 1) * B]
        y, acc = O.i8_conv(xs.astype(np.int32), s_x, z_x, w[s].astype(np.int32), s_w, z_w, bias, s_out, z_out, stride, pad, 1, relu, act_bits=7)
        if with_res:
            y = O.qadd(y.transpose(0, 2, 3, 1), s_out, z_out, res_q[s * B:(s + 1) * B].astype(np.int32).transpose(0, 2, 3, 1), s_res, z_res,
                       s_add, z_add, 0, 255).transpose(0, 3, 1, 2)
            y = np.clip(y, max(0, z_add), 127)              # add -> clamp -> ReLU -> clamp (models_bbb.py:178-182) == add_relu floor
        want.append(y)
        want_acc.append(acc)
    want = np.concatenate(want, 0)

    # ---- CUDA
    xm = _to_map(ops, xq, s_x, z_x, phase_split=(stride == 2))
    wb = ops.i8_p16_block_weights(torch.as_tensor(w).cuda(), xm.C_pad, stride)
    if out_split:
        out = ops.P16Map.empty(n * B, N, Ho // 2 + 1, Ho // 2 + 1, 4, "cuda")
    else:
        out = ops.P16Map.empty(n * B, N, Ho + 1, Ho + 1, 1, "cuda")
    rm = _to_map(ops, res_q, s_res, z_res) if with_res else None
    Hp_o = Ho + 1
    acc_dump = torch.zeros((n * B * Hp_o * Hp_o, N), dtype=torch.int32, device="cuda")
    ops.i8_conv_p16_forward(xm, wb, n, N, k, k, stride, torch.as_tensor(bias).cuda(), s_w, z_w, s_out, z_out, relu, 7, out, residual=rm,
                            add_qp=(s_add, z_add) if with_res else None, add_relu=True, x_shared=x_shared, out_phase_split=out_split,
                            acc_dump=acc_dump)
    torch.cuda.synchronize()
    got = out.to_quint8().cpu().numpy().astype(np.int32)
    assert got.shape == want.shape
    assert np.array_equal(got, want), "requantised outputs differ: %d of %d" % (int((got != want).sum()), got.size)
    acc_got = acc_dump.reshape(n * B, Hp_o, Hp_o, N)[:, 1:, 1:, :].permute(0, 3, 1, 2).cpu().numpy()
    assert np.array_equal(acc_got, np.concatenate(want_acc, 0)), "int32 accumulators differ"
    # the padding the next layer relies on: borders, padded channel planes and the tail stay zero
    if not out_split:
        full = out.buf[:, :n * B * Hp_o * Hp_o].reshape(-1, n * B, Hp_o, Hp_o, 16)
        assert int(full[:, :, 0].abs().sum()) == 0 and int(full[:, :, :, 0].abs().sum()) == 0
        planes = full.permute(1, 2, 3, 0, 4).reshape(n * B, Hp_o, Hp_o, -1)
        assert int(planes[..., N:].abs().sum()) == 0
    assert int(out.buf[:, out.phases * n * B * out.Hp * out.Wp:].abs().sum()) == 0


def test_i8_planar_layout_round_trip_and_avgpool():
    ops = _ops()
    rng = np.random.default_rng(5)
    q = _rand_q(rng, (6, 192, 4, 4))
    m = _to_map(ops, q, 0.1, 37)
    assert np.array_equal(m.to_quint8().cpu().numpy(), q)
    pooled = ops.i8_p16_avgpool(m, 7).cpu().numpy().astype(np.int32)
    want = O.i8_avgpool(q.astype(np.int32), 37, 4, act_bits=7).reshape(6, 192)
    assert np.array_equal(pooled, want)
    q2 = _rand_q(rng, (4, 24, 8, 8))
    m2 = _to_map(ops, q2, 0.1, 11, phase_split=True)
    assert np.array_equal(m2.to_quint8().cpu().numpy(), q2)


def test_i8_planar_blocked_weights_layout():
    ops = _ops()
    rng = np.random.default_rng(6)
    n, N, C, k = 2, 48, 24, 3
    w = rng.integers(-128, 128, size=(n, N, C, k, k), dtype=np.int64).astype(np.int8)
    wb = ops.i8_p16_block_weights(torch.as_tensor(w).cuda(), 32, 1).cpu().numpy()
    n_pad = 64
    blk = wb.reshape(n, 1, 9, 2, n_pad, 16)                         # [sample][cb][tap][chunk][row][16]
    for t in range(9):
        r, s = divmod(t, 3)
        got = blk[:, 0, t].transpose(0, 2, 1, 3).reshape(n, n_pad, 32)      # [sample][row][channel]
        assert np.array_equal(got[:, :N, :C], w[:, :, :, r, s])
        assert np.all(got[:, N, :C] == 1) and np.all(got[:, N, C:] == 0)    # the ones row covers the real channels only
        assert np.all(got[:, :N, C:] == 0) and np.all(got[:, N + 1:] == 0)
