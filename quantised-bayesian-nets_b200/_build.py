"""In-tree build of csrc/*.cu -> csrc/libqbn.so for sm_100a (nvcc cross-compiles without a GPU).

The shared object is git-ignored but travels to the GPU box with the snapshot; the product
never JIT-compiles and never falls back when the library is missing (see _lib.py)."""
import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libqbn.so")
SOURCES = ["core.cu", "gemm_fp32.cu", "i8_conv.cu", "i8_p16.cu", "reduce.cu", "umma_conv.cu", "umma_conv_s1.cu", "umma_conv_p4.cu", "umma_wgrad.cu", "umma_wgrad_p4.cu", "lrt_p4.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]
if os.environ.get("QBN_TUNING"):          # kernel-tuning builds only: environment knobs / cycle accounting (scripts/p4_sweep.sh)
    NVCC_FLAGS.append("-DQBN_TUNING")


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _digest():
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)) + ["../../include/qbn.h"]:
        p = os.path.join(CSRC, name)
        if name.endswith((".cu", ".cuh", ".h")) and os.path.isfile(p):
            h.update(name.encode())
            h.update(open(p, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_lib(force=False, verbose=False):
    stamp = os.path.join(CSRC, ".build_stamp")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out.decode()))
        if verbose and out:
            print(out.decode())
    cmd = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s" % r.stdout.decode())
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build_lib(force=True, verbose=True))
