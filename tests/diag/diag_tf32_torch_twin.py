"""Context for DESIGN.md §6: how far do the parameter gradients of the UNMODIFIED reference network move when torch itself switches
its convolutions from fp32 to TF32 (cuDNN allow_tf32), same seeds, same batch?  If torch's own twin shows the same ~10 % relative-L2
spread as this library's TF32 vs fp32 modes, the spread is a property of the network (ReLU masks), not of the kernels.
Usage (GPU box): python tests/diag/diag_tf32_torch_twin.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    net = bench._reference_model()
    if net is None:
        print("reference modules unavailable")
        return
    net = net.cuda().train()
    g = torch.Generator().manual_seed(2)
    x = torch.randn(256, 3, 32, 32, generator=g).cuda()
    t = torch.randint(0, 10, (256,), generator=g).cuda()
    grads = {}
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        net.zero_grad(set_to_none=True)
        torch.manual_seed(11)
        out = net(x)
        kl = net.get_kl_divergence()
        loss = torch.nn.functional.nll_loss(torch.log(out + 1e-8), t) + 0.01 * kl / (256 * 176)
        loss.backward()
        grads[tf32] = {k: p.grad.detach().double().clone() for k, p in net.named_parameters() if p.grad is not None}
        print("reference network on this GPU, cuDNN allow_tf32=%s: loss %.6f" % (tf32, float(loss.detach())))
    rel = torch.tensor([float((grads[True][k] - grads[False][k]).norm() / (grads[False][k].norm() + 1e-30)) for k in grads[False]])
    print("relative L2 difference of the %d parameter gradients, torch TF32 vs torch fp32, same seeds: median %.3e  p90 %.3e  max %.3e"
          % (len(rel), float(rel.median()), float(torch.quantile(rel, 0.9)), float(rel.max())))


if __name__ == "__main__":
    main()
