"""Fixed cost of one evaluation pass (config 5, B=256): graph-replay time of MCEngine.predict_sum for several sample counts,
the a + b*S fit, and the kernel list of the S=13 pass (the per-rank share at 8 GPUs).  Usage: python scripts/fixed_cost.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import __graft_entry__ as ge
    ge.build()
    import numpy as np
    from qbn_b200 import noise, synthetic, zoo
    from qbn_b200.mc import MCEngine
    model = zoo.resnet_from_params(synthetic.ResNetBBBParams(seed=1)).cuda().eval()
    noise.manual_seed(1)
    x = torch.randn(256, 3, 32, 32, generator=torch.Generator().manual_seed(3)).cuda()
    eng = MCEngine(model, chunk=int(os.environ.get("QBN_CHUNK", "50")), lanes=int(os.environ.get("QBN_LANES", "1")),
                   sample_ahead=os.environ.get("QBN_AHEAD", "0") == "1")
    pts = []
    for S in (1, 2, 4, 6, 12, 13, 25, 50, 100):
        for _ in range(3):
            eng.predict_sum(x, S)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            eng.predict_sum(x, S)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        pts.append((S, ms))
        print("S=%3d  %.3f ms  (%.4f ms/sample)" % (S, ms, ms / S))
    A = np.array([[1.0, s] for s, _ in pts if s <= 50])
    y = np.array([m for s, m in pts if s <= 50])
    a, b = np.linalg.lstsq(A, y, rcond=None)[0]
    print("fit over S<=50: %.3f ms + %.4f ms * S" % (a, b))
    from torch.profiler import ProfilerActivity, profile
    for S in (13,):
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            eng.predict_sum(x, S)
            torch.cuda.synchronize()
        evs = [(ev.time_range.start, ev.time_range.end, ev.name, (ev.device_time if hasattr(ev, "device_time") else ev.cuda_time))
               for ev in prof.events() if ev.device_type is not None and "cuda" in str(ev.device_type).lower()]
        evs.sort()
        busy = sum(e[3] for e in evs)
        span = evs[-1][1] - evs[0][0]
        print("S=%d: %d launches, kernel time %.3f ms, span first..last %.3f ms (gaps %.3f ms)" % (S, len(evs), busy / 1e3, span / 1e3, (span - busy) / 1e3))
        rows = {}
        for _, _, n, us in evs:
            r = rows.setdefault(n, [0, 0.0])
            r[0] += 1
            r[1] += us
        for n, (c, us) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:25]:
            print("  %8.1f us %5.1f %% %4d x  %s" % (us, 100 * us / busy, c, n[:120]))
        print("launch order (us): " + " ".join("%s:%.0f" % (n.split("::")[-1].split("(")[0].split("<")[0].replace("_kernel", "")[:18], us) for _, _, n, us in evs))


if __name__ == "__main__":
    main()
