"""CPU, world_size 2, gloo: the multi-GPU host logic of qbn_b200/dist.py (sample sharding + the single
allreduce of probability sums; data-parallel gradient allreduce after the per-replica NaN scrub)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from qbn_b200 import dist as qd
    S, B, K = 11, 6, 10
    g = torch.Generator().manual_seed(0)
    probs = torch.softmax(torch.randn(S, B, K, generator=g), -1)          # "sample s" output, identical on every rank
    start, count = qd.shard_range(S, rank, world)
    psum = probs[start:start + count].sum(0)
    qd.allreduce_prob_sums(psum)
    ok1 = torch.allclose(psum / S, probs.mean(0), atol=1e-6)
    # regression heads: three running sums
    mu = torch.randn(S, B, generator=g)
    var = torch.rand(S, B, generator=g) + 0.1
    m, v = qd.reduce_regression(mu[start:start + count].sum(0), (mu[start:start + count] ** 2).sum(0), var[start:start + count].sum(0), S)
    ok2 = torch.allclose(m, mu.mean(0), atol=1e-5) and torch.allclose(v, mu.var(0) + var.mean(0), atol=1e-4)
    # data-parallel gradients: NaN scrub per replica, then one flat allreduce (mean)
    lin = torch.nn.Linear(4, 3)
    torch.manual_seed(1)
    for p in lin.parameters():
        p.grad = torch.full_like(p, float(rank + 1))
    if rank == 1:
        lin.weight.grad[0, 0] = float("nan")
    qd.scrub_nan_grads(lin.parameters())
    qd.allreduce_gradients(list(lin.parameters()), average=True)
    exp = torch.full_like(lin.weight, 1.5)
    exp[0, 0] = 0.5                                                       # (1 + 0) / 2: the NaN was zeroed before the reduce
    ok3 = torch.allclose(lin.weight.grad, exp) and torch.allclose(lin.bias.grad, torch.full_like(lin.bias, 1.5))
    qd.broadcast_parameters(lin)
    # ShardedMCPredictor over a stub engine: more ranks than samples (the idle rank contributes zeros and still enters the
    # collective), regression through the three running sums, and SGHMC member sharding (member index = sample index)
    class _Cls:
        regression, n_classes, model = False, K, None

        def predict_sum(self, x, count, sample0=0):
            return probs[sample0:sample0 + count].sum(0).clone()

    class _Reg:
        regression, model = True, None

        def predict_sum(self, x, count, sample0=0):
            return mu[sample0:sample0 + count].clone(), var[sample0:sample0 + count].clone()
    xb = torch.zeros(B, 3)
    ok4 = torch.allclose(qd.ShardedMCPredictor(_Cls()).predict(xb, 1), probs[0], atol=1e-6)           # rank 1 owns no sample
    ok4 = ok4 and torch.allclose(qd.ShardedMCPredictor(_Cls()).predict(xb, S), probs.mean(0), atol=1e-6)
    m2, v2 = qd.ShardedMCPredictor(_Reg()).predict(xb, S)
    ok5 = torch.allclose(m2, mu.mean(0), atol=1e-5) and torch.allclose(v2, mu.var(0) + var.mean(0), atol=1e-4)
    m3, v3 = qd.ShardedMCPredictor(_Reg()).predict(xb, 1)                                             # S=1 on 2 ranks: var of one draw = 0
    ok5 = ok5 and torch.allclose(m3, mu[0], atol=1e-6)
    # an engine that takes a unit window (MCEngine.supports_window): the balanced (sample, image) split of shard_units —
    # S=11 samples x 6 images over 2 ranks = 5.5 samples each; rank 0's last sample and rank 1's first are the same one
    class _Win:
        regression, n_classes, model, supports_window = False, K, None, True
        seen = []

        def predict_sum(self, x, count, sample0=0, window=None):
            first, end = window if window is not None else (0, B)
            self.seen.append((sample0, count, first, end))
            mask = torch.ones(count, B, 1)
            mask[0, :first] = 0
            mask[count - 1, end:] = 0
            return (probs[sample0:sample0 + count] * mask).sum(0)
    w = _Win()
    ok6 = torch.allclose(qd.ShardedMCPredictor(w).predict(xb, S), probs.mean(0), atol=1e-6)
    ok6 = ok6 and w.seen[-1] == ((0, 6, 0, 3) if rank == 0 else (5, 6, 3, 6))
    ret[rank] = bool(ok1 and ok2 and ok3 and ok4 and ok5 and ok6)
    dist.destroy_process_group()


def test_shard_range_partitions():
    from qbn_b200.dist import shard_range
    for S in (1, 7, 100):
        for w in (1, 2, 3, 8):
            parts = [shard_range(S, r, w) for r in range(w)]
            assert sum(c for _, c in parts) == S
            assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    assert shard_range(100, 0, 8) == (0, 13) and shard_range(100, 7, 8) == (88, 12)     # ideal speed-up 100/13 = 7.69x


def test_shard_units_partitions():
    """Every (sample, image) unit belongs to exactly one rank; the ranks' unit counts differ by at most one; with world | samples the
    split is shard_range's."""
    from qbn_b200.dist import shard_range, shard_units
    for S, B, W in ((100, 256, 8), (100, 256, 3), (5, 7, 8), (1, 4, 3), (3, 2, 16), (12, 5, 4)):
        seen, sizes = set(), []
        for r in range(W):
            s0, n, first, end = shard_units(S, B, r, W)
            mine = 0
            for i in range(n):
                for im in range(first if i == 0 else 0, end if i == n - 1 else B):
                    assert (s0 + i, im) not in seen
                    seen.add((s0 + i, im))
                    mine += 1
            sizes.append(mine)
        assert len(seen) == S * B and max(sizes) - min(sizes) <= 1
        if S % W == 0:
            assert all(shard_units(S, B, r, W) == (*shard_range(S, r, W), 0, B) for r in range(W))
    assert shard_units(100, 256, 0, 8) == (0, 13, 0, 128) and shard_units(100, 256, 1, 8) == (12, 13, 128, 256)   # 12.5 samples each


def test_shard_units_property():
    """hypothesis: for any (samples, batch, world) the windows tile the (sample, image) grid exactly once, in rank order, balanced."""
    from hypothesis import given, settings, strategies as st
    from qbn_b200.dist import shard_units

    @settings(max_examples=300, deadline=None)
    @given(st.integers(1, 40), st.integers(1, 33), st.integers(1, 24))
    def check(S, B, W):
        nxt, sizes = 0, []
        for r in range(W):
            s0, n, first, end = shard_units(S, B, r, W)
            if n == 0:
                sizes.append(0)
                continue
            assert 0 <= first < B and 0 < end <= B and (n > 1 or first < end)
            lo, hi = s0 * B + first, (s0 + n - 1) * B + end          # unit range [lo, hi) in sample-major order
            assert lo == nxt and hi > lo
            nxt = hi
            sizes.append(hi - lo)
        assert nxt == S * B and max(sizes) - min(sizes) <= 1
    check()


def test_world2_gloo():
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
