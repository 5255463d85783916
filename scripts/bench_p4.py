"""Back-to-back launches of qbn_conv_p4_fwd for every ResNet layer shape (GPU-side time between two CUDA
events).  Usage: python scripts/bench_p4.py [samples]"""
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
for (C, H, N, k, stride, res, split) in [(24, 32, 24, 3, 1, False, False), (24, 32, 24, 3, 1, True, False), (24, 32, 24, 3, 1, True, True),
                                         (24, 16, 48, 3, 2, False, False), (24, 16, 48, 1, 2, False, False),
                                         (48, 16, 48, 3, 1, True, False), (48, 8, 96, 3, 2, False, False), (48, 8, 96, 1, 2, False, False),
                                         (96, 8, 96, 3, 1, True, False), (96, 4, 192, 3, 2, False, False), (96, 4, 192, 1, 2, False, False),
                                         (192, 4, 192, 3, 1, True, False)]:
    # H = OUTPUT resolution
    Hp = H + 1
    phases = 4 if stride == 2 else 1
    x = ops.P4Map.empty(S * B, C, Hp, Hp, (1, 1), phases, "cuda")
    x.buf[:, :phases * S * B * Hp * Hp].copy_(rnd(torch.randn(C // 4, phases * S * B * Hp * Hp, 4, device="cuda")))
    w = rnd(torch.randn(S, ops.p4_weight_floats(C, N, k, k, stride), device="cuda") * 0.05)
    r = None
    if res:
        r = ops.P4Map.empty(S * B, N, Hp, Hp, (1, 1), 1, "cuda")
        r.buf.normal_()
    if split:
        out = ops.P4Map.empty(S * B, N, H // 2 + 1, H // 2 + 1, (1, 1), 4, "cuda")
    else:
        out = ops.P4Map.empty(S * B, N, Hp, Hp, (1, 1), 1, "cuda")
    sc = torch.rand(N, device="cuda") + 0.5; sh = torch.randn(N, device="cuda")
    ms = timeit(lambda: ops.conv_p4_forward(x, w, S, N, k, k, stride, sc, sh, r, True, ops.QBN_FLAG_OUT_ROUND_TF32, False, out, split))
    fl = 2.0 * S * B * H * H * N * k * k * C
    by = 4.0 * (x.buf.numel() + out.buf.numel() + (r.buf.numel() if res else 0))
    print("p4  C%3d out%2dx%-2d N%3d k%d s%d res=%d split=%d : %7.1f us  %6.1f TF/s  %6.0f GB/s" % (C, H, H, N, k, stride, res, split, ms * 1e3, fl / ms / 1e9, by / ms / 1e6))
for (C2, N, H) in [(24, 48, 16), (48, 96, 8), (96, 192, 4)]:
    Hp = H + 1
    y = ops.P4Map.empty(S * B, N, Hp, Hp, (1, 1), 1, "cuda")
    y.buf[:, :S * B * Hp * Hp].copy_(rnd(torch.randn(N // 4, S * B * Hp * Hp, 4, device="cuda")))
    x2 = ops.P4Map.empty(S * B, C2, Hp, Hp, (1, 1), 4, "cuda")
    x2.buf.copy_(rnd(torch.randn_like(x2.buf)))
    cb2 = ops.p4_shortcut_block_channels(N, C2)
    nfl = ops.p4_weight_floats(N, N, 3, 3, 1) + (C2 // cb2) * (cb2 // 4) * ((N + 15) // 16 * 16) * 4
    w = rnd(torch.randn(S, nfl, device="cuda") * 0.05)
    out = ops.P4Map.empty(S * B, N, Hp, Hp, (1, 1), 1, "cuda")
    sh = torch.randn(N, device="cuda")
    ms = timeit(lambda: ops.conv_p4_shortcut_forward(y, w, x2, S, N, 3, 3, None, sh, True, ops.QBN_FLAG_OUT_ROUND_TF32, out))
    print("p4  fused shortcut C%3d->N%3d out%2dx%-2d (3x3 + 1x1/2 of C%d) : %7.1f us" % (N, N, H, H, C2, ms * 1e3))
n = 1571592
mu = torch.randn(n, device="cuda"); sg = torch.rand(n, device="cuda")
ms = timeit(lambda: ops.sample_weights(mu, sg, S, None, 1, 2, 0, True))
print("sample_weights (canonical) all layers (%d x %d): %7.1f us" % (S, n, ms * 1e3))
tot = 0.0
for (C, N, cnt) in [(24, 24, 4), (24, 48, 1), (48, 48, 3), (48, 96, 1), (96, 96, 3), (96, 192, 1), (192, 192, 3)]:
    mu = torch.randn(N * 9 * C, device="cuda"); sg = torch.rand(N * 9 * C, device="cuda")
    mb, sb = ops.p4_block_weights(mu, N, C, 9)[0], ops.p4_block_weights(sg, N, C, 9)[0]
    out = torch.empty(S, mb.numel(), device="cuda")
    ms = timeit(lambda: ops.sample_weights_blocked(mb, sb, N, C, 9, S, None, 1, 2, 0, True, out))
    tot += ms * cnt
    print("sample_weights_blocked C%3d N%3d: %7.1f us" % (C, N, ms * 1e3))
print("blocked sampler, all 3x3 layers: %.1f us" % (tot * 1e3))
