"""GPU kernel durations of one LRT training step (config 4) from torch.profiler (CUPTI): every kernel of a replay of the graphed step,
grouped by name.  Usage: python scripts/profile_train_kernels.py [B]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    import __graft_entry__ as ge
    ge.build()
    from qbn_b200 import config, losses, noise, synthetic, zoo
    from qbn_b200 import dist as qdist
    config.set_math_mode("tf32")
    config.set_pdl(os.environ.get("QBN_PDL", "0") == "1")
    model = zoo.resnet_from_params(synthetic.ResNetBBBParams(seed=1)).cuda().train()
    noise.manual_seed(1)
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-3, capturable=True, fused=True)
    crit = losses.LOSS_FACTORY["classification"](zoo.Args(loss_multiplier=1.0), "batch")
    g = torch.Generator().manual_seed(5)
    x, t = torch.randn(B, 3, 32, 32, generator=g).cuda(), torch.randint(0, 10, (B,), generator=g).cuda()
    step = qdist.GraphedTrainStep(model, crit, opt, x, t, 176, 45000, gamma=0.01, warmup=3)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    print("graphed step: %.3f ms" % (e0.elapsed_time(e1) / 10))
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    rows = {}
    for ev in prof.events():
        if ev.device_type is not None and "cuda" in str(ev.device_type).lower():
            n = ev.name
            r = rows.setdefault(n, [0, 0.0])
            r[0] += 1
            r[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    ours = [(ev.time_range.start, ev.name, (ev.device_time if hasattr(ev, "device_time") else ev.cuda_time)) for ev in prof.events()
            if ev.device_type is not None and "cuda" in str(ev.device_type).lower()
            and any(k in ev.name for k in ("umma_", "w32_", "p4_stage", "lrt_p4", "lrt_stage"))]
    ours.sort()
    print("our kernels in launch order (us):")
    print("  " + " ".join("%s:%.0f" % (n.split("::")[-1].split("(")[0].replace("_kernel", "")[:22], us) for _, n, us in ours))
    tot = sum(v[1] for v in rows.values())
    print("kernel time of one replay: %.3f ms in %d launches" % (tot / 1e3, sum(v[0] for v in rows.values())))
    for n, (c, us) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:40]:
        print("  %8.1f us %5.1f %% %4d x  %s" % (us, 100 * us / tot, c, n[:110] if "Functor" not in n else n[:400]))


if __name__ == "__main__":
    main()
