"""Diagnostics of qbn_lrt_wgrad_p4 / dgrad on a tiny problem (tuning builds: QBN_WG_V descriptor variants).
Usage: [QBN_WG_V=n] python scripts/dbg/wgrad_diag.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import __graft_entry__ as ge
ge.build()
from qbn_b200 import ops

torch.manual_seed(0)
for (B, C, H, N, k, stride, pad) in ((1, 8, 6, 8, 3, 1, 1), (2, 24, 16, 24, 3, 1, 1), (2, 24, 16, 48, 3, 2, 1)):
    x = torch.randn(B, C, H, H).cuda()
    mu = (torch.randn(N, C, k, k) / (C * k * k) ** 0.5).cuda()
    rho = torch.empty(N, C, k, k).uniform_(-5, -2).cuda()
    d = ops.make_desc(B, H, H, C, N, k, k, stride, pad, 1)
    Ho = d.Ho
    eps = torch.randn(B, N, Ho, Ho).cuda()
    go = torch.randn(B, N, Ho, Ho).cuda()
    xc = ops.nhwc(x)
    out, std, x_p4, xsq_p4 = ops.lrt_p4_forward(xc, mu, rho, False, None, d, ops.nhwc(eps))
    dx, dmu_p, dsig2_p = ops.lrt_p4_backward(xc, x_p4, xsq_p4, std, ops.nhwc(eps), mu, rho, False, ops.nhwc(go), d, (0, 0, 0), True)
    torch.cuda.synchronize()
    # references in torch (fp32 on the GPU)
    dmu_ref = torch.nn.grad.conv2d_weight(x, mu.shape, go, stride, pad)
    dv = go * eps / (2 * std)
    dsig2_ref = torch.nn.grad.conv2d_weight(x * x, mu.shape, dv, stride, pad)
    sig2 = torch.nn.functional.softplus(rho) ** 2
    dx_ref = torch.nn.grad.conv2d_input(x.shape, mu, go, stride, pad) + 2 * x * torch.nn.grad.conv2d_input(x.shape, sig2, dv, stride, pad)
    got_mu = dmu_p.view(N, k, k, C).permute(0, 3, 1, 2)
    got_s2 = dsig2_p.view(N, k, k, C).permute(0, 3, 1, 2)

    def rel(a, b):
        return float((a - b).norm() / (b.norm() + 1e-30))
    print("shape", (B, C, H, N, k, stride, pad), "V=%s" % os.environ.get("QBN_WG_V", "0"))
    print("  dx    rel err %.3e   max|got| %.3e max|ref| %.3e" % (rel(dx, dx_ref), float(dx.abs().max()), float(dx_ref.abs().max())))
    print("  dmu   rel err %.3e   max|got| %.3e max|ref| %.3e  nonzero %d / %d" % (rel(got_mu, dmu_ref), float(got_mu.abs().max()), float(dmu_ref.abs().max()),
                                                                                    int((got_mu != 0).sum()), got_mu.numel()))
    print("  dsig2 rel err %.3e   max|got| %.3e max|ref| %.3e" % (rel(got_s2, dsig2_ref), float(got_s2.abs().max()), float(dsig2_ref.abs().max())))
    # one-hot probe: g = 1 at (pixel (2,3), channel 1), x = 1 at (pixel (2,3), channel 2) -> dmu[1][centre tap][2] = 1
    if stride == 1:
        x1 = torch.zeros_like(x); x1[0, 2, 2, 3] = 1.0
        g1 = torch.zeros_like(go); g1[0, 1, 2, 3] = 1.0
        xc1 = ops.nhwc(x1)
        o1, s1, xp1, xq1 = ops.lrt_p4_forward(xc1, mu, rho, False, None, d, ops.nhwc(eps))
        _, m1, _ = ops.lrt_p4_backward(xc1, xp1, xq1, s1, ops.nhwc(eps), mu, rho, False, ops.nhwc(g1), d, (0, 0, 0), False)
        nz = torch.nonzero(m1.view(N, k * k, C))
        print("  one-hot probe: expected [[1, %d, 2]] got %s values %s" % (k * k // 2, nz.tolist()[:8], m1[m1 != 0][:8].tolist()))
