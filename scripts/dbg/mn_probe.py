"""Which shared-memory layouts does tcgen05.mma kind::tf32 accept for MN-major operands?  Hypothesis test on one MMA (M=128, N=32, K=8..).
Builds A / B images under a layout hypothesis, runs the probe, compares D with A @ B."""
import ctypes, itertools, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "quantised-bayesian-nets_b200", "csrc")
SO = os.path.join(HERE, "_mn_probe.so")
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-I", CSRC,
                       os.path.join(HERE, "mn_probe.cu"), "-o", SO])
lib = ctypes.CDLL(SO)
lib.mn_probe.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_int,
                         ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]

M, N = 128, 32


def desc(lbo, sbo, layout_type, base_offset=0):
    d = ((lbo >> 4) & 0x3FFF) << 16 | ((sbo >> 4) & 0x3FFF) << 32 | 1 << 46 | (base_offset & 7) << 49 | (layout_type & 7) << 61
    return d


def idesc(a_mn, b_mn, n=N):
    return (1 << 4) | (2 << 7) | (2 << 10) | (int(a_mn) << 15) | (int(b_mn) << 16) | ((n >> 3) << 17) | ((128 >> 4) << 24)


def tf32(x):
    i = x.view(np.int32).copy()
    i = (i + 0x1000) & ~0x1FFF
    return i.view(np.float32)


def run(a_img, b_img, adesc, bdesc, a_off, b_off, idsc, n_kstep=1, a_kadv=0, b_kadv=0):
    a = torch.from_numpy(a_img.view(np.uint8)).cuda()
    b = torch.from_numpy(b_img.view(np.uint8)).cuda()
    D = torch.full((M, N), float("nan"), device="cuda")
    rc = lib.mn_probe(a.data_ptr(), b.data_ptr(), a.numel(), b.numel(), adesc, bdesc, a_off, b_off, idsc, N, n_kstep, a_kadv, b_kadv, D.data_ptr())
    torch.cuda.synchronize()
    assert rc == 0, rc
    return D.cpu().numpy()


def build_mn_sw(mat, K, lbo, sbo, xor_mode, row0=0, rows_alloc=None, group=4, chunk=32):
    """mat [MN][K] -> image bytes.  Rows of 128 B = 32 MN-contiguous floats; K row k at (k // group) * sbo + (k % group) * 128 + row0 * 128;
    MN block b (32 values) at b * lbo; 32-byte chunk index XORed with key(xor_mode)."""
    MN = mat.shape[0]
    size = 65536
    img = np.zeros(size // 4, dtype=np.float32)
    for mn in range(MN):
        for k in range(K):
            row_byte = (mn // 32) * lbo + (k // group) * sbo + (k % group) * 128 + row0 * 128
            p = (mn % 32) * 4
            c, within = p // chunk, p % chunk
            if xor_mode == "abs":
                key = (row_byte >> 7) & (128 // chunk - 1)
            elif xor_mode == "rel":
                key = k % (128 // chunk)
            else:
                key = 0
            off = row_byte + ((c ^ key) * chunk) + within
            img[off // 4] = mat[mn, k]
    return img


def build_k_major(mat, K, lbo, sbo):
    """the K-major no-swizzle layout the product kernels use (known good): mat [MN][K]"""
    img = np.zeros(65536 // 4, dtype=np.float32)
    for mn in range(mat.shape[0]):
        for k in range(K):
            off = (mn // 8) * sbo + (mn % 8) * 16 + (k // 4) * lbo + (k % 4) * 4
            img[off // 4] = mat[mn, k]
    return img


rng = np.random.default_rng(0)
K = 8
A = tf32(rng.standard_normal((M, K)).astype(np.float32))
B = tf32(rng.standard_normal((N, K)).astype(np.float32))
ref = A @ B.T


def report(name, D):
    err = np.abs(D - ref).max()
    print("%-100s max|D| %9.3e  err %9.3e  %s" % (name, np.nanmax(np.abs(D)), err, "MATCH" if err < 1e-3 * np.abs(ref).max() else ""))


# 0. sanity: both K-major (the product layout)
a_img = build_k_major(A, K, lbo=128 * 16, sbo=128)
b_img = build_k_major(B, K, lbo=32 * 16, sbo=128)
report("K-major A, K-major B (sanity)", run(a_img, b_img, desc(128 * 16, 128, 0), desc(32 * 16, 128, 0), 0, 0, idesc(0, 0)))

# 1. MN-major A (layout type t), K-major B — and the mirror
for lt in (1, 0, 2, 4, 6):
    for xor_mode in ("abs", "none"):
        for (lbo, sbo, swap) in ((4096, 512, False), (4096, 512, True)):
            a_img = build_mn_sw(A, K, lbo, sbo, xor_mode)
            dl, ds = (sbo, lbo) if swap else (lbo, sbo)
            D = run(a_img, b_img, desc(dl, ds, lt), desc(32 * 16, 128, 0), 0, 0, idesc(1, 0))
            report("MN-major A: layout_type %d xor %-4s LBO %d SBO %d" % (lt, xor_mode, dl, ds), D)
# chunk granularity variants of the XOR for layout type 1 (16-byte chunks keyed by row % 8: the plain 128B swizzle) and type 2
for lt, chunk in ((2, 16), (1, 16), (2, 32)):
    a_img = build_mn_sw(A, K, 4096, 1024 if chunk == 16 else 512, "abs", group=8 if chunk == 16 else 4, chunk=chunk)
    D = run(a_img, b_img, desc(4096, 1024 if chunk == 16 else 512, lt), desc(32 * 16, 128, 0), 0, 0, idesc(1, 0))
    report("MN-major A: layout_type %d, %d-byte chunks keyed by row %% %d" % (lt, chunk, 128 // chunk), D)

# 2. both MN-major, layout type 1
a_img = build_mn_sw(A, K, 4096, 512, "abs")
bm_img = build_mn_sw(B, K, 4096, 512, "abs")
report("MN-major A and B (type 1, abs xor)", run(a_img, bm_img, desc(4096, 512, 1), desc(4096, 512, 1), 0, 0, idesc(1, 1)))

# 3. start address shifted by r rows (tap shift): data built with the ABSOLUTE key, base_offset 0 / (r & 3) / (r & 7)
for r in (1, 2, 3, 4, 5, 35):
    a_img = build_mn_sw(A, K, 8192, 512, "abs", row0=r)
    bm_img = build_mn_sw(B, K, 8192, 512, "abs", row0=r)
    for bo in (0, r & 3, r & 7):
        D = run(a_img, bm_img, desc(8192, 512, 1, bo), desc(8192, 512, 1, bo), r * 128, r * 128, idesc(1, 1))
        report("both MN-major, start shifted by %d rows, absolute-keyed data, base_offset %d" % (r, bo), D)
    a_img = build_mn_sw(A, K, 8192, 512, "rel", row0=r)
    bm_img = build_mn_sw(B, K, 8192, 512, "rel", row0=r)
    D = run(a_img, bm_img, desc(8192, 512, 1, 0), desc(8192, 512, 1, 0), r * 128, r * 128, idesc(1, 1))
    report("both MN-major, start shifted by %d rows, start-relative-keyed data, base_offset 0" % r, D)

# 4. K = 16 as two MMAs advancing both start addresses by 8 rows (1024 B)
K2 = 16
A2 = tf32(rng.standard_normal((M, K2)).astype(np.float32))
B2 = tf32(rng.standard_normal((N, K2)).astype(np.float32))
ref = A2 @ B2.T
a_img = build_mn_sw(A2, K2, 8192, 512, "abs", row0=3)
bm_img = build_mn_sw(B2, K2, 8192, 512, "abs", row0=3)
report("K = 16: two MMAs, +1024 B per step, start at row 3", run(a_img, bm_img, desc(8192, 512, 1), desc(8192, 512, 1), 3 * 128, 3 * 128, idesc(1, 1), 2, 1024, 1024))
