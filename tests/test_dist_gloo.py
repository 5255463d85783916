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
    ret[rank] = bool(ok1 and ok2 and ok3 and ok4 and ok5)
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


def test_world2_gloo():
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
