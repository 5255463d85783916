import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
from qbn_b200 import _lib, ops
names = ["clock overhead", "tcgen05.fence::after", "fence.proxy.async", "commit+wait round trip (idle pipe)", "dependent MMA issue", "dependent MMA issue+complete",
         "misaligned-A MMA issue", "misaligned-A MMA issue+complete", "4-accumulator MMA issue", "4-accumulator issue+complete", "LDTM.x16 + wait", "commit issue only"]
for n_cols in (32, 64, 128):
    for reps in (8, 64):
        out = torch.zeros(16, dtype=torch.int64, device="cuda")
        _lib.call("qbn_ubench_tcgen05", ctypes.c_void_p(out.data_ptr()), n_cols, reps, ops._stream())
        torch.cuda.synchronize()
        o = out.cpu().tolist()
        print("N=%d reps=%d: " % (n_cols, reps) + "; ".join("%s=%d" % (n, v) for n, v in zip(names, o)))
        print("   4 warps issuing concurrently (issue, issue+complete per MMA): " + ", ".join("w%d=(%d,%d)" % (w, o[12 + w] >> 32, o[12 + w] & 0xffffffff) for w in range(4)))

print("multi-CTA / multi-issuer overlap (cycles per MMA seen by each issuer of CTA 0: issue, issue+complete)")
for n_cols in (32, 64, 128):
    for (k, w) in [(1, 1), (1, 2), (1, 3), (1, 4), (2, 1), (3, 1), (4, 1), (2, 2)]:
        if w * n_cols * k > 512:
            continue
        out = torch.zeros(4, dtype=torch.int64, device="cuda")
        _lib.call("qbn_ubench_tcgen05_multi", ctypes.c_void_p(out.data_ptr()), n_cols, 64, k, w, ops._stream())
        torch.cuda.synchronize()
        o = out.cpu().tolist()
        per = [(v >> 32, v & 0xffffffff) for v in o[:w]]
        print("N=%d  %d CTA/SM x %d issuer warps: " % (n_cols, k, w) + ", ".join("(%d,%d)" % p for p in per) + "   -> SM-wide %.0f cycles/MMA" % (per[0][1] / (k * w)))
