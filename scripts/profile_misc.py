"""One launch of every NON-conv hot-path kernel (and of the LRT / int8 tcgen05 contractions) at BASELINE sizes, for an ncu pass:
   ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
Sizes: ResNet-18 BBB, B=256, one 10-sample chunk."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as ge
ge.build()
from qbn_b200 import ops
S, B = 10, 256
g = torch.Generator().manual_seed(1)
torch.cuda.synchronize()
# A4: canonical sampler over all 1.57 M weights x 10 samples; blocked sampler on the widest layer
n = 1571592
mu, sg = torch.randn(n, generator=g).cuda(), torch.rand(n, generator=g).cuda()
ops.sample_weights(mu, sg, S, None, 1, 2, 0, True)
C = N = 192
mub = ops.p4_block_weights(torch.randn(N * 9 * C, generator=g).cuda(), N, C, 9)[0]
sgb = ops.p4_block_weights(torch.rand(N * 9 * C, generator=g).cuda(), N, C, 9)[0]
ops.sample_weights_blocked(mub, sgb, N, C, 9, S, None, 1, 2, 0, True)
# A5: KL + gradient over all weights
mu_p = mu.clone().requires_grad_(True); rho_p = (sg - 5.0).clone().requires_grad_(True)
ops.kl_divergence(mu_p, rho_p, 0.05).backward()
# A7: observer + fake-quant of a layer-1 activation (B=256, 24x32x32) and of the weights
act = torch.randn(B, 24, 32, 32, generator=g).cuda()
fq = ops.FakeQuantState(0, 127)
ops.fake_quantize(act, fq, True)
fqw = ops.FakeQuantState(-128, 127)
ops.fake_quantize(mu, fqw, True)
# A8: dropout on the same activation
ops.dropout_forward(ops.nhwc(act), 0.15, None, (1, 2, 3))
# A9 / A10: softmax accumulation over S x B x 10 logits, classification metrics
logits = torch.randn(100, B, 10, generator=g).cuda()
ps = ops.softmax_accumulate(logits, None)
acc = torch.zeros(4 + 30, device="cuda")
ops.cls_metrics_accumulate(ps / 100.0, torch.randint(0, 10, (B,), generator=g).cuda(), acc)
# A1/A2: LRT forward, tcgen05 dual accumulator (layer-1 shape, B=256)
x = ops.nhwc(torch.randn(B, 24, 32, 32, generator=g).cuda())
wmu, wrho = torch.randn(24, 24, 3, 3, generator=g).cuda() * 0.07, torch.full((24, 24, 3, 3), -5.0).cuda()
p = ops.weight_prep(wmu, wrho, False, None, want=("mu", "sigma2"), round_tf32=True)
d = ops.make_desc(B, 32, 32, 24, 24, 3, 3, 1, 1, 1)
ops.lrt_forward(x, p["mu"], p["sigma2"], None, d, None, (5, 6, 7), ops.QBN_MATH_TF32)
# A6: int8 sampling + kind::i8 conv + requantisation (layer-1 shape, 10 samples)
rng = np.random.default_rng(0)
xq = torch.as_tensor(rng.integers(0, 128, (S * B, 24, 32, 32)).astype(np.uint8)).cuda().contiguous(memory_format=torch.channels_last)
wq = torch.as_tensor(rng.integers(-128, 128, (S, 24 * 9 * 24)).astype(np.int8)).cuda()
ops.i8_conv_forward(xq, 0.021, 17, wq, 0.0037, -3, d, None, 0.09, 40, 1, 7, S, False, False, False, path=2)
pp = ops.I8SampleParams()
pp.s_mu, pp.z_mu, pp.s_sigma, pp.z_sigma = 0.004, 3, 0.0005, -128
pp.s_eps, pp.z_eps, pp.s_mul, pp.z_mul, pp.s_add, pp.z_add = 3.0 / 127, 0, 0.0006, 0, 0.0045, 2
pp.w_min, pp.w_max, pp.n_vec = -128, 127, -1
muq = torch.as_tensor(rng.integers(-128, 128, n).astype(np.int8)).cuda()
sgq = torch.as_tensor(rng.integers(-128, 128, n).astype(np.int8)).cuda()
ops.i8_sample_weights(muq, sgq, pp, S, None, 1, 2, 0)
torch.cuda.synchronize()
print("ok")

# ---- achieved HBM bandwidth of the bandwidth-bound kernels at LARGE sizes (CUDA events, 10 back-to-back launches, buffers > L2)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
if os.environ.get("QBN_MISC_BW", "1") == "1" and "ncu" not in os.environ.get("NV_NSIGHT_INJECTION_PORT_BASE", "") and not os.environ.get("NV_COMPUTE_PROFILER_PERFWORKS_DIR"):
    big = torch.randn(S * B, 24, 32, 32, generator=torch.Generator(device="cuda").manual_seed(1), device="cuda")   # 252 MB
    nb = big.numel() * 4
    rows = []
    S2 = 100
    wbig = torch.empty((S2, n), device="cuda")
    ms = timeit(lambda: _raw_sample(mu, sg, wbig, S2))  if False else None
    ms = timeit(lambda: ops.sample_weights(mu, sg, S2, None, 1, 2, 0, True))
    rows.append(("sample_weights (S=100 x 1.57 M weights: 8 B read once (L2), 4 B written per weight)", ms, 4.0 * S2 * n + 8.0 * n))
    fqb = ops.FakeQuantState(0, 127)
    ms = timeit(lambda: ops.fake_quantize(big, fqb, True))
    rows.append(("fake-quant observer + quantise (252 MB activation: 4 B + 4 B read, 4 B written per element)", ms, 3.0 * nb))
    ms = timeit(lambda: ops.fake_quantize(big, fqb, False))
    rows.append(("fake-quant quantise only (4 B read, 4 B written per element)", ms, 2.0 * nb))
    bigc = ops.nhwc(big)
    mask = (torch.rand(S * B, 24, device="cuda") < 0.85).float()
    ms = timeit(lambda: ops.dropout_forward(bigc, 0.15, mask))
    rows.append(("MC-Dropout apply (4 B read, 4 B written per element)", ms, 2.0 * nb))
    q = ops.quantize_u8(big, 0.05, 3)
    ms = timeit(lambda: ops.quantize_u8(big, 0.05, 3))
    rows.append(("quantize fp32 -> u8 (4 B read, 1 B written per element)", ms, 1.25 * nb))
    lg = torch.randn(100, 4096, 10, device="cuda")
    ms = timeit(lambda: ops.softmax_accumulate(lg, None))
    rows.append(("softmax_accumulate S=100 x B=4096 x 10 (16 MB read)", ms, lg.numel() * 4.0))
    mu_l, rho_l = torch.randn(64 * n // 16, device="cuda").requires_grad_(True), torch.randn(64 * n // 16, device="cuda").requires_grad_(True)
    from qbn_b200 import _lib
    klo, dmu, drho = torch.zeros(1, device="cuda"), torch.zeros_like(mu_l), torch.zeros_like(rho_l)
    P_ = lambda t: __import__("ctypes").c_void_p(t.data_ptr())
    ms = timeit(lambda: _lib.call("qbn_kl_fwd_bwd", P_(mu_l), P_(rho_l), mu_l.numel(), 0.05, P_(klo), P_(dmu), P_(drho), 1.0, ops._stream()))
    rows.append(("KL value + gradient accumulate, %d weights (8 B read, 8 B read-modify-written per weight)" % mu_l.numel(), ms, 24.0 * mu_l.numel()))
    peak = 6550.1
    print("%-100s %9s %9s %7s" % ("kernel (algorithmic bytes)", "us", "GB/s", "of peak"))
    for name, ms, by in rows:
        print("%-100s %9.1f %9.0f %6.1f%%" % (name, ms * 1e3, by / ms / 1e6, 100.0 * by / ms / 1e6 / peak))
