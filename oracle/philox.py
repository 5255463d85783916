"""TEST INFRASTRUCTURE (oracle) — bit-exact numpy restatement of the device RNG in
quantised-bayesian-nets_b200/csrc/common.cuh (Philox4x32-10, Salmon et al. SC'11; the same
generator family torch.cuda uses for normal_()/bernoulli_() behind linear.py:36-37,44-45,
conv.py:28-29,34-35 and dropout.py:21-30).  The reference has no RNG of its own to restate: its
draws come from torch's global generator, which the parity tests replay by injection instead.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over the counter arrays (uint32).  Returns four uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint32).copy()
    c1 = np.asarray(c1, dtype=np.uint32).copy()
    c2 = np.asarray(c2, dtype=np.uint32).copy()
    c3 = np.asarray(c3, dtype=np.uint32).copy()
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & MASK32).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & MASK32).astype(np.uint32)
            n0 = hi1 ^ c1 ^ k0
            n2 = hi0 ^ c3 ^ k1
            c0, c1, c2, c3 = n0, lo1, n2, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def philox_u32(n, seed, stream_a, stream_b):
    """Stream element i = word (i & 3) of Philox(counter = i >> 2, stream_a, stream_b; key = seed)."""
    n4 = (n + 3) // 4
    ctr = np.arange(n4, dtype=np.uint64)
    c0 = (ctr & MASK32).astype(np.uint32)
    c1 = (ctr >> np.uint64(32)).astype(np.uint32)
    c2 = np.full(n4, stream_a, dtype=np.uint32)
    c3 = np.full(n4, stream_b, dtype=np.uint32)
    x, y, z, w = philox4x32_10(c0, c1, c2, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return np.stack([x, y, z, w], axis=1).reshape(-1)[:n]


def u01(x):
    """24-bit uniform strictly inside (0,1): (x>>8)*2^-24 + 2^-25 (exact in fp32)."""
    return ((x >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24) + np.float32(2.0 ** -25)).astype(np.float32)


def philox_normal(n, seed, stream_a, stream_b):
    """Box-Muller on word pairs (x,y)->(z0,z1), (z,w)->(z2,z3); fp32 like the device code (the
    device uses logf/sqrtf/sincospif, so values agree to a few ulp, not bit for bit)."""
    n4 = (n + 3) // 4 * 4
    u = philox_u32(n4, seed, stream_a, stream_b).reshape(-1, 4)
    out = np.empty((u.shape[0], 4), dtype=np.float32)
    for j in (0, 2):
        u1 = u01(u[:, j]).astype(np.float64)
        u2 = u01(u[:, j + 1]).astype(np.float64)
        r = np.sqrt(-2.0 * np.log(u1))
        out[:, j] = (r * np.cos(2.0 * np.pi * u2)).astype(np.float32)
        out[:, j + 1] = (r * np.sin(2.0 * np.pi * u2)).astype(np.float32)
    return out.reshape(-1)[:n]


def philox_bernoulli(n, keep_prob, seed, stream_a, stream_b):
    """1.0 with probability keep_prob: u01(word) < keep_prob (dropout.py:21-30 bernoulli_(1-p))."""
    u = u01(philox_u32(n, seed, stream_a, stream_b))
    return (u < np.float32(keep_prob)).astype(np.float32)
