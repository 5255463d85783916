"""Per-entry-point timing of one LRT training step (config 4): CUDA events around every libqbn call, torch glue = the rest.
Usage: python scripts/profile_train.py [tf32|fp32]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    math = sys.argv[1] if len(sys.argv) > 1 else "tf32"
    import __graft_entry__ as ge
    ge.build()
    from qbn_b200 import _lib, config, losses, noise, synthetic, zoo
    from qbn_b200 import dist as qdist
    config.set_math_mode(math)
    model = zoo.resnet_from_params(synthetic.ResNetBBBParams(seed=1)).cuda().train()
    noise.manual_seed(1)
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-3)
    step = qdist.DPTrainStep(model, losses.LOSS_FACTORY["classification"](zoo.Args(loss_multiplier=1.0), "batch"), opt, gamma=0.01, check_nan_loss=False)
    g = torch.Generator().manual_seed(5)
    x, t = torch.randn(256, 3, 32, 32, generator=g).cuda(), torch.randint(0, 10, (256,), generator=g).cuda()
    for _ in range(3):
        step(x, t, 176, 45000)
    rows, real = [], _lib.call

    def timed(name, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = real(name, *a)
        e1.record()
        rows.append((name, e0, e1))
        return out
    _lib.call = timed
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    step(x, t, 176, 45000)
    t1.record()
    torch.cuda.synchronize()
    _lib.call = real
    total = t0.elapsed_time(t1)
    per, cnt = {}, {}
    for n, a, b in rows:
        per[n] = per.get(n, 0.0) + a.elapsed_time(b)
        cnt[n] = cnt.get(n, 0) + 1
    print("LRT training step, B=256, %s: %.3f ms (with per-call events)" % (math, total))
    for n in sorted(per, key=per.get, reverse=True):
        print("  %-28s %4d calls %9.3f ms %5.1f %%" % (n, cnt[n], per[n], 100 * per[n] / total))
    print("  %-28s            %9.3f ms %5.1f %%" % ("torch (BN, ReLU, add, pool, Adam, glue)", total - sum(per.values()), 100 * (total - sum(per.values())) / total))
    fw = [(a.elapsed_time(b)) for n, a, b in rows if n == "qbn_lrt_fwd"]
    bw = [(a.elapsed_time(b)) for n, a, b in rows if n == "qbn_lrt_bwd"]
    print("  qbn_lrt_fwd per layer:", " ".join("%.3f" % v for v in fw))
    print("  qbn_lrt_bwd per layer:", " ".join("%.3f" % v for v in bw))


if __name__ == "__main__":
    main()
