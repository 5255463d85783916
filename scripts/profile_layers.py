"""Per-launch timing table of one MC chunk (CUDA events around every libqbn call of the engine)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
from qbn_b200 import synthetic as O
from qbn_b200 import mc, noise, zoo, ops
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 10
P = O.ResNetBBBParams(seed=1)
model = zoo.resnet_from_params(P).cuda().eval()
noise.manual_seed(1)
eng = mc.MCEngine(model, math_mode="tf32", chunk=chunk, use_graph=False)
x = torch.randn(256, 3, 32, 32, generator=torch.Generator().manual_seed(2)).cuda()
rows = []
def wrap(name, fn, describe):
    def f(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(*a, **k); e1.record()
        rows.append((name, describe(a, k, out), e0, e1))
        return out
    return f
def d_conv(a, k, out):
    d, n = a[2], a[3]
    fl = 2.0 * n * d.B * d.Ho * d.Wo * d.N * d.R * d.S * d.C
    by = 4.0 * (a[0].numel() + out.numel() + (a[8].numel() if a[8] is not None else 0))
    return dict(shape="B%d %dx%d C%d->N%d k%d s%d" % (d.B, d.H, d.W, d.C, d.N, d.R, d.stride_h), flops=fl, bytes=by, mode=a[12])
def d_s1(a, k, out):
    x_, w, n, N, R, S_ = a[:6]
    H, W = x_.shape[2] - (R - 1), x_.shape[3] - (S_ - 1)
    fl = 2.0 * x_.shape[0] * H * W * N * R * S_ * x_.shape[1]
    res = a[8] if len(a) > 8 else None
    by = 4.0 * (x_.numel() + out.numel() + (res.numel() if res is not None else 0))
    return dict(shape="SB%d %dx%d C%d->N%d k%d" % (x_.shape[0], H, W, x_.shape[1], N, R), flops=fl, bytes=by, mode=1)
def d_samp(a, k, out): return dict(shape="n=%d S=%d" % (a[0].numel(), a[2]), flops=0, bytes=4.0 * out.numel(), mode=-1)
def d_other(a, k, out): return dict(shape="", flops=0, bytes=0, mode=-1)
mc.ops.conv_forward = wrap("conv_v1", ops.conv_forward, d_conv)
mc.ops.conv_s1_forward = wrap("conv_s1", ops.conv_s1_forward, d_s1)
mc.ops.sample_weights = wrap("sample_w", ops.sample_weights, d_samp)
mc.ops.avgpool_all = wrap("avgpool", ops.avgpool_all, d_other)
def d_p4(a, k, out):
    x_, w, n, N, R, S_ = a[:6]
    stride = a[6] if len(a) > 6 else 1
    H, W = x_.Hp - 1, x_.Wp - 1
    fl = 2.0 * x_.n_img * H * W * N * R * S_ * x_.C
    res = a[9] if len(a) > 9 else None
    by = 4.0 * (x_.buf.numel() * (0.25 if (stride == 2 and R == 1) else 1.0) + out.buf.numel() + (res.buf.numel() if res is not None else 0))
    return dict(shape="SB%d out%dx%d C%d->N%d k%d s%d%s%s" % (x_.n_img, H, W, x_.C, N, R, stride, " +res" if res is not None else "", " split" if out.phases == 4 else ""),
                flops=fl, bytes=by, mode=1)
mc.ops.conv_p4_forward = wrap("conv_p4", ops.conv_p4_forward, d_p4)
def d_p4sc(a, k, out):
    x_, w, x2 = a[:3]
    n, N, R, S_ = a[3:7]
    H, W = x_.Hp - 1, x_.Wp - 1
    fl = 2.0 * x_.n_img * H * W * N * (R * S_ * x_.C + x2.C)
    by = 4.0 * (x_.buf.numel() + 0.25 * x2.buf.numel() + out.buf.numel())
    return dict(shape="SB%d out%dx%d C%d->N%d k%d + fused 1x1/2 of C%d" % (x_.n_img, H, W, x_.C, N, R, x2.C), flops=fl, bytes=by, mode=1)
mc.ops.conv_p4_shortcut_forward = wrap("conv_p4", ops.conv_p4_shortcut_forward, d_p4sc)
mc.ops.sample_weights_blocked = wrap("sample_wb", ops.sample_weights_blocked, d_other)
mc.ops.avgpool_p4 = wrap("avgpool_p4", ops.avgpool_p4, d_other)
mc.ops.softmax_accumulate = wrap("softmax_acc", ops.softmax_accumulate, d_other)
for _ in range(2):
    rows.clear()
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(); eng.predict_sum(x, chunk, 0); t1.record()
    torch.cuda.synchronize()
print("chunk of %d samples: total %.3f ms (GPU events)" % (chunk, t0.elapsed_time(t1)))
tot = {}
for name, d, e0, e1 in rows:
    ms = e0.elapsed_time(e1)
    tot[name] = tot.get(name, 0) + ms
    if name.startswith("conv"):
        print("%-8s %-34s %8.3f ms  %7.1f TF/s  %7.0f GB/s %s" % (name, d["shape"], ms, d["flops"] / ms / 1e9, d["bytes"] / ms / 1e6, "fp32" if d["mode"] == 0 else ""))
print({k: round(v, 3) for k, v in tot.items()}, "sum", round(sum(tot.values()), 3))
