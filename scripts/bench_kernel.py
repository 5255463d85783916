"""Back-to-back launches of one kernel configuration between two CUDA events (GPU-side time, not
CPU launch pace).  Usage: python scripts/bench_kernel.py [samples]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
from qbn_b200 import ops
S = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B = 256
reps = 20
def timeit(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def rnd(t):
    ti = t.view(torch.int32); ti.add_(0x1000).bitwise_and_(~0x1FFF); return t
print("S=%d B=%d" % (S, B))
for (C, H, N, res) in [(24, 32, 24, False), (24, 32, 24, True), (48, 16, 48, False), (48, 16, 48, True), (96, 8, 96, True), (192, 4, 192, True)]:
    x = rnd(torch.randn(S * B, C, H + 2, H + 2, device="cuda").contiguous(memory_format=torch.channels_last))
    w = rnd(torch.randn(S, N * 9 * C, device="cuda") * 0.05)
    r = torch.randn(S * B, N, H + 2, H + 2, device="cuda").contiguous(memory_format=torch.channels_last) if res else None
    out = torch.empty(S * B, N, H + 2, H + 2, device="cuda").contiguous(memory_format=torch.channels_last)
    sc = torch.rand(N, device="cuda") + 0.5; sh = torch.randn(N, device="cuda")
    ms = timeit(lambda: ops.conv_s1_forward(x, w, S, N, 3, 3, sc, sh, r, True, ops.QBN_FLAG_OUT_ROUND_TF32, False, out))
    fl = 2.0 * S * B * H * H * N * 9 * C
    by = 4.0 * (x.numel() + out.numel() + (r.numel() if res else 0))
    print("s1  C%3d %2dx%-2d N%3d res=%d : %7.1f us  %6.1f TF/s  %6.0f GB/s" % (C, H, H, N, res, ms * 1e3, fl / ms / 1e9, by / ms / 1e6))
# v1: first layer (shared input, stacked), stride-2 3x3 and 1x1 from bordered input
x0 = torch.randn(B, 4, 32, 32, device="cuda").contiguous(memory_format=torch.channels_last)
w0 = rnd(torch.randn(S, 24 * 9 * 4, device="cuda") * 0.1)
d0 = ops.make_desc(B, 32, 32, 4, 24, 3, 3, 1, 1, 1); d0.out_pad_h = d0.out_pad_w = 1
o0 = torch.zeros(S * B, 24, 34, 34, device="cuda").contiguous(memory_format=torch.channels_last)
ms = timeit(lambda: ops.conv_forward(x0, w0, d0, S, True, False, None, None, None, True, None, 1.0, ops.QBN_MATH_TF32, o0, ops.QBN_FLAG_OUT_ROUND_TF32))
print("v1  conv0 C4->24 shared-x       : %7.1f us  %6.0f GB/s (write)" % (ms * 1e3, 4.0 * o0.numel() / ms / 1e6))
for (C, H, N, k) in [(24, 32, 48, 3), (24, 32, 48, 1), (48, 16, 96, 3), (48, 16, 96, 1), (96, 8, 192, 3), (96, 8, 192, 1)]:
    x = rnd(torch.randn(S * B, C, H + 2, H + 2, device="cuda").contiguous(memory_format=torch.channels_last))
    w = rnd(torch.randn(S, N * k * k * C, device="cuda") * 0.05)
    d = ops.make_desc(B, H + 2, H + 2, C, N, k, k, 2, (0 if k == 3 else -1), 1); d.out_pad_h = d.out_pad_w = 1
    out = torch.zeros(S * B, N, d.Ho + 2, d.Wo + 2, device="cuda").contiguous(memory_format=torch.channels_last)
    ms = timeit(lambda: ops.conv_forward(x, w, d, S, False, False, None, None, None, False, None, 1.0, ops.QBN_MATH_TF32, out, ops.QBN_FLAG_OUT_ROUND_TF32 | ops.QBN_FLAG_A_TF32_READY))
    fl = 2.0 * S * B * d.Ho * d.Wo * N * k * k * C
    print("v1  C%3d %2dx%-2d N%3d k%d s2      : %7.1f us  %6.1f TF/s  %6.0f GB/s" % (C, H, H, N, k, ms * 1e3, fl / ms / 1e9, 4.0 * (x.numel() + out.numel()) / ms / 1e6))
n = 1571592
mu = torch.randn(n, device="cuda"); sg = torch.rand(n, device="cuda")
ms = timeit(lambda: ops.sample_weights(mu, sg, S, None, 1, 2, 0, True))
print("sample_weights all layers (%d x %d): %7.1f us  %6.0f GB/s written" % (S, n, ms * 1e3, 4.0 * S * n / ms / 1e6))
