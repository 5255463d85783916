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
