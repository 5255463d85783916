"""Measured error distribution of the TF32 paths against the fp32 mode of the same library (rtol 1e-5 vs the reference) on the
BASELINE-size workload: (1) ResNet-18 BBB eval, B=256, S=10 identical Philox draws — per-probability absolute and relative error of
the MC-averaged softmax; (2) one LRT training step, B=256, identical noise — relative L2 error of every parameter gradient.
Usage: python scripts/tf32_error_stats.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pct(t, qs=(0.5, 0.9, 0.99, 0.999, 1.0)):
    t = t.flatten().double().cpu()
    return " ".join("p%g=%.3e" % (100 * q, float(torch.quantile(t, q))) for q in qs)


def main():
    import __graft_entry__ as ge
    ge.build()
    from qbn_b200 import config, losses, mc, noise, synthetic, zoo
    g = torch.Generator().manual_seed(2)
    x = torch.randn(256, 3, 32, 32, generator=g).cuda()
    t = torch.randint(0, 10, (256,), generator=g).cuda()
    # ---- eval
    model = zoo.resnet_from_params(synthetic.ResNetBBBParams(seed=1)).cuda().eval()
    probs = {}
    for mode in ("fp32", "tf32"):
        noise.manual_seed(7)
        eng = mc.MCEngine(model, math_mode=mode, chunk=10)
        probs[mode] = (eng.predict_sum(x, 10, sample0=0) / 10).double()
    d = (probs["tf32"] - probs["fp32"]).abs()
    print("eval, B=256, S=10, MC-averaged probabilities (2560 values), TF32 vs fp32 mode of the same draws:")
    print("  |dp|            :", pct(d))
    print("  |dp| / p        :", pct(d / probs["fp32"].clamp_min(1e-12)))
    print("  |dp| / max_k p  :", pct(d / probs["fp32"].max(dim=1, keepdim=True).values))
    print("  argmax flips    : %d of 256" % int((probs["tf32"].argmax(1) != probs["fp32"].argmax(1)).sum()))
    # ---- training step gradients
    crit = losses.LOSS_FACTORY["classification"](zoo.Args(loss_multiplier=1.0), "batch")
    grads = {}
    first_id = noise._state["next_layer_id"]
    for mode in ("fp32", "tf32"):
        config.set_math_mode(mode)
        noise._state["next_layer_id"] = first_id
        m = zoo.resnet_from_params(synthetic.ResNetBBBParams(seed=1)).cuda().train()
        noise.manual_seed(11)
        out = m(x)
        loss, _, _ = crit(out, t, m.get_kl_divergence(), 0.01, 176, 45000)
        loss.backward()
        grads[mode] = {k: p.grad.detach().double() for k, p in m.named_parameters() if p.grad is not None}
        print("training step (%s): loss %.6f" % (mode, float(loss)))
    rel = {k: float((grads["tf32"][k] - grads["fp32"][k]).norm() / (grads["fp32"][k].norm() + 1e-30)) for k in grads["fp32"]}
    vals = torch.tensor(list(rel.values()))
    print("relative L2 error of the %d parameter gradients, TF32 vs fp32 mode, same noise:" % len(rel))
    print("  ", pct(vals, (0.5, 0.9, 1.0)))
    worst = sorted(rel.items(), key=lambda kv: -kv[1])[:5]
    print("   worst:", ", ".join("%s %.2e" % kv for kv in worst))


if __name__ == "__main__":
    main()
