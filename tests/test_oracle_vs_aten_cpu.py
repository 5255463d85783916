"""CPU (pytest -m "not gpu"): the integer half of the oracle against the arithmetic the reference actually executes — ATen's
QuantizedCPU kernels and FBGEMM, which are third-party to the reference (pinned torch==1.7.1 in its requirements.txt:54,
torch 2.11 installed here) but present as a binary.  The golden fixtures pin a handful of fixed tensors; this
file sweeps random scales, zero points, shapes and saturating inputs, so the restated rounding sequences are pinned over
their whole input range (bit-exact everywhere)."""
import numpy as np
import pytest
import torch

import oracle.qbn_oracle as O


@pytest.fixture(autouse=True)
def _one_thread_fbgemm():
    # ATen splits a quantised elementwise op into one vector body + scalar tail PER THREAD CHUNK; one thread makes the
    # split what the oracle models (n_vec = n // 64 * 64).  The reference fixes the engine the same way (quant_utils.py:118).
    old_threads, old_engine = torch.get_num_threads(), torch.backends.quantized.engine
    torch.set_num_threads(1)
    torch.backends.quantized.engine = "fbgemm"
    yield
    torch.set_num_threads(old_threads)
    torch.backends.quantized.engine = old_engine


def _qparams(rng, signed):
    scale = float(np.float32(10.0 ** rng.uniform(-3.5, -0.5)))
    zp = int(rng.integers(-128, 128) if signed else rng.integers(0, 256))
    return scale, zp


def _qtensor(ints, scale, zp, signed):
    t = torch.as_tensor(np.asarray(ints, np.int8 if signed else np.uint8))
    return torch._make_per_tensor_quantized_tensor(t, scale, zp)


@pytest.mark.parametrize("seed", range(6))
def test_quantize_mul_add_sweep(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.choice([1, 63, 64, 65, 1000, 4096 + 17]))
    # quantize_per_tensor, signed and unsigned, with inputs far outside the representable range
    x = (rng.standard_normal(n) * 10.0 ** rng.uniform(-2, 2)).astype(np.float32)
    for signed in (True, False):
        s, z = _qparams(rng, signed)
        lo, hi = (-128, 127) if signed else (0, 255)
        ref = torch.quantize_per_tensor(torch.as_tensor(x), s, z, torch.qint8 if signed else torch.quint8).int_repr().numpy()
        assert np.array_equal(O.quantize(x, s, z, lo, hi), ref), (signed, s, z)
    # quantized::mul and quantized::add on qint8 (the weight-sampling recipe, linear_q.py:86-92)
    a, b = rng.integers(-128, 128, n), rng.integers(-128, 128, n)
    (sa, za), (sb, zb), (so, zo) = _qparams(rng, True), _qparams(rng, True), _qparams(rng, True)
    qa, qb = _qtensor(a, sa, za, True), _qtensor(b, sb, zb, True)
    assert np.array_equal(O.qmul(a, sa, za, b, sb, zb, so, zo), torch.ops.quantized.mul(qa, qb, so, zo).int_repr().numpy())
    assert np.array_equal(O.qadd(a, sa, za, b, sb, zb, so, zo), torch.ops.quantized.add(qa, qb, so, zo).int_repr().numpy())
    # quantized::add / add_relu on quint8 (the residual add, src/utils.py:49-55)
    a, b = rng.integers(0, 256, n), rng.integers(0, 256, n)
    (sa, za), (sb, zb), (so, zo) = _qparams(rng, False), _qparams(rng, False), _qparams(rng, False)
    qa, qb = _qtensor(a, sa, za, False), _qtensor(b, sb, zb, False)
    assert np.array_equal(O.qadd(a, sa, za, b, sb, zb, so, zo, 0, 255), torch.ops.quantized.add(qa, qb, so, zo).int_repr().numpy())
    assert np.array_equal(O.qadd(a, sa, za, b, sb, zb, so, zo, 0, 255, relu=True),
                          torch.ops.quantized.add_relu(qa, qb, so, zo).int_repr().numpy())


@pytest.mark.parametrize("seed", range(4))
def test_linear_and_conv_requantisation_sweep(seed):
    rng = np.random.default_rng(100 + seed)
    (s_x, z_x), (s_w, z_w), (s_o, z_o) = _qparams(rng, False), _qparams(rng, True), _qparams(rng, False)
    s_o *= 30.0                                                   # keep most outputs inside [0, 255]; the rest saturates
    B, K, N = int(rng.integers(1, 9)), int(rng.integers(1, 70)), int(rng.integers(1, 40))
    x, w = rng.integers(0, 256, (B, K)), rng.integers(-128, 128, (N, K))
    bias = rng.standard_normal(N).astype(np.float32)
    for relu in (False, True):
        for b in (None, bias):
            packed = torch.ops.quantized.linear_prepack(_qtensor(w, s_w, z_w, True), None if b is None else torch.as_tensor(b))
            op = torch.ops.quantized.linear_relu if relu else torch.ops.quantized.linear
            ref = op(_qtensor(x, s_x, z_x, False), packed, s_o, z_o).int_repr().numpy()
            got, _ = O.i8_linear(x, s_x, z_x, w, s_w, z_w, b, s_o, z_o, relu, act_bits=8)
            assert np.array_equal(got, ref), (relu, b is not None)
    C, Nc, H, k = int(rng.integers(1, 9)), int(rng.integers(1, 12)), int(rng.integers(5, 11)), int(rng.choice([1, 3, 5]))
    stride, pad = int(rng.choice([1, 2])), int(rng.choice([0, k // 2]))
    x, w = rng.integers(0, 256, (2, C, H, H)), rng.integers(-128, 128, (Nc, C, k, k))
    bias = rng.standard_normal(Nc).astype(np.float32)
    for relu in (False, True):
        packed = torch.ops.quantized.conv2d_prepack(_qtensor(w, s_w, z_w, True), torch.as_tensor(bias), [stride] * 2, [pad] * 2, [1, 1], 1)
        op = torch.ops.quantized.conv2d_relu if relu else torch.ops.quantized.conv2d
        ref = op(_qtensor(x, s_x, z_x, False), packed, s_o, z_o).int_repr().numpy()
        got, _ = O.i8_conv(x, s_x, z_x, w, s_w, z_w, bias, s_o, z_o, stride, pad, 1, relu, act_bits=8)
        assert np.array_equal(got, ref), (relu, C, Nc, H, k, stride, pad)


@pytest.mark.parametrize("seed", range(4))
def test_relu_avgpool_clamp_sweep(seed):
    rng = np.random.default_rng(200 + seed)
    s, z = _qparams(rng, False)
    k = int(rng.choice([2, 4]))
    x = rng.integers(0, 256, (3, int(rng.integers(1, 9)), 2 * k, 3 * k))
    q = _qtensor(x, s, z, False).contiguous(memory_format=torch.channels_last)
    assert np.array_equal(O.i8_relu(x, z, act_bits=8), torch.relu(q).int_repr().numpy())
    assert np.array_equal(O.i8_avgpool(x, z, k, act_bits=8), torch.nn.functional.avg_pool2d(q, k).int_repr().numpy())
    # clamp_activation (src/utils.py:25-30): a float clamp of the quantised tensor == an integer clamp, qparams unchanged
    for bits in (7, 4):
        lo, hi = O.UINT_BOUNDS[bits]
        ref = torch.clamp(q, (lo - z) * s, (hi - z) * s)
        assert ref.q_scale() == q.q_scale() and ref.q_zero_point() == z
        assert np.array_equal(np.clip(x, lo, hi), ref.int_repr().numpy())
        assert np.array_equal(O.i8_relu(x, 0, act_bits=bits), np.clip(x, lo, hi))


@pytest.mark.parametrize("seed", range(4))
def test_fake_quantize_sweep(seed):
    rng = np.random.default_rng(300 + seed)
    x = (rng.standard_normal(5000) * 10.0 ** rng.uniform(-2, 1)).astype(np.float32)
    for qmin, qmax in ((0, 127), (-128, 127), (-8, 7)):
        s = float(np.float32(10.0 ** rng.uniform(-3, -1)))
        z = int(rng.integers(qmin, qmax + 1))
        ref = torch.fake_quantize_per_tensor_affine(torch.as_tensor(x), s, z, qmin, qmax).numpy()
        got, mask = O.fake_quant(x, s, z, qmin, qmax)
        assert np.array_equal(got, ref), (qmin, qmax, s, z)
        t = torch.as_tensor(x).requires_grad_(True)
        torch.fake_quantize_per_tensor_affine(t, s, z, qmin, qmax).sum().backward()
        assert np.array_equal(mask, t.grad.numpy() != 0)              # straight-through gradient only where not clamped


@pytest.mark.parametrize("seed", range(4))
def test_add_relu_is_a_floor_at_the_output_zero_point(seed):
    """What `ops.i8_add(relu=True)` relies on: quantized::add_relu == quantized::add followed by max(q, zero_point)."""
    rng = np.random.default_rng(400 + seed)
    n = 4096
    a, b = rng.integers(0, 256, n), rng.integers(0, 256, n)
    (sa, za), (sb, zb), (so, zo) = _qparams(rng, False), _qparams(rng, False), _qparams(rng, False)
    ref = torch.ops.quantized.add_relu(_qtensor(a, sa, za, False), _qtensor(b, sb, zb, False), so, zo).int_repr().numpy()
    assert np.array_equal(np.maximum(O.qadd(a, sa, za, b, sb, zb, so, zo, 0, 255), zo), ref)
