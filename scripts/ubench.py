import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import subprocess
import torch
import __graft_entry__ as ge
ge.build()
from qbn_b200 import _lib, ops
# the micro-benchmark is not part of the product library: built here, on demand, against libqbn.so (error plumbing, SM count)
HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "quantised-bayesian-nets_b200", "csrc")
UB = os.path.join(HERE, "_libqbn_ubench.so")
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-I", CSRC,
                       os.path.join(HERE, "ubench.cu"), "-o", UB, "-L", CSRC, "-lqbn", "-Xlinker", "-rpath=" + CSRC])
_lib.load()
_ub = ctypes.CDLL(UB)
def _ub_call(name, *a):
    _lib.check(getattr(_ub, name)(*a), name)
names = ["clock overhead", "tcgen05.fence::after", "fence.proxy.async", "commit+wait round trip (idle pipe)", "dependent MMA issue", "dependent MMA issue+complete",
         "misaligned-A MMA issue", "misaligned-A MMA issue+complete", "4-accumulator MMA issue", "4-accumulator issue+complete", "LDTM.x16 + wait", "commit issue only"]
for n_cols in (32, 64, 128):
    for reps in (8, 64):
        out = torch.zeros(16, dtype=torch.int64, device="cuda")
        _ub_call("qbn_ubench_tcgen05", ctypes.c_void_p(out.data_ptr()), n_cols, reps, ops._stream())
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
        _ub_call("qbn_ubench_tcgen05_multi", ctypes.c_void_p(out.data_ptr()), n_cols, 64, k, w, ops._stream())
        torch.cuda.synchronize()
        o = out.cpu().tolist()
        per = [(v >> 32, v & 0xffffffff) for v in o[:w]]
        print("N=%d  %d CTA/SM x %d issuer warps: " % (n_cols, k, w) + ", ".join("(%d,%d)" % p for p in per) + "   -> SM-wide %.0f cycles/MMA" % (per[0][1] / (k * w)))
