"""Multi-GPU plumbing the reference does not have: one process per GPU,
torch.distributed (NCCL on B200s, gloo in CPU tests).

  * MC-sample sharding (inference): rank r owns a contiguous range of GLOBAL sample indices; noise
    is keyed by the global index, so the union over ranks equals the single-GPU draw set.  Each
    rank accumulates sum_s softmax locally; ONE allreduce(SUM) of the [B,K] buffer per batch.
  * Data-parallel LRT training: parameters replicated, batch sharded, ONE flat allreduce(SUM) of
    all gradients per step (12.6 MB for the ResNet: latency-bound, so a single bucket), after the
    per-replica NaN-grad scrub of trainer.py:105-107.  The KL term is parameter-only: every rank
    computes it and scales it with the GLOBAL batch (losses.py:24), it is not reduced twice.
"""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous split of `total` samples: (start, count) for `rank`; counts differ by at most 1."""
    base, rem = divmod(total, world)
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def shard_units(samples, batch, rank, world):
    """Balanced split of the samples x batch (sample, image) units in sample-major order: rank r owns units
    [r*U/W, (r+1)*U/W), U = samples * batch.  Returns (sample0, n_samples, first_img, end_img): the rank runs samples
    [sample0, sample0 + n_samples); of the first one it owns images [first_img, batch) only, of the last images [0, end_img) only
    (both restrictions on the same sample when n_samples == 1).  With world | samples this is shard_range; otherwise no rank
    carries a whole extra sample (S=100 on 8 GPUs: 12.5 samples each instead of 13 / 12)."""
    total = samples * batch
    u0, u1 = rank * total // world, (rank + 1) * total // world
    if u1 <= u0:
        return 0, 0, 0, batch
    s_first, first_img = divmod(u0, batch)
    s_last, end_img = divmod(u1 - 1, batch)
    return s_first, s_last - s_first + 1, first_img, end_img + 1


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class ShardedMCPredictor:
    """predict(x, S): every rank gets the full batch x; returns the same result on every rank — p-bar [B, K] for
    classification, (mean, variance) of experiments/utils.py:349-353 for regression."""

    def __init__(self, engine):
        self.engine = engine

    def predict(self, x, samples):
        rank, ws = world()
        start, count = shard_range(samples, rank, ws)
        if self.engine.regression:
            # the three running sums of the reference's formula; a rank without samples contributes zeros and still
            # enters the collective (raising here would leave the other ranks hanging in all_reduce)
            if count > 0:
                mu, var = self.engine.predict_sum(x, count, sample0=start)          # [count, B] each
                sums = (mu.sum(0), (mu * mu).sum(0), var.sum(0))
            else:
                z = torch.zeros(x.shape[0], dtype=torch.float32, device=x.device)
                sums = (z, z.clone(), z.clone())
            return reduce_regression(*sums, samples)
        out = self.local_sum(x, samples, rank, ws, None)
        if ws > 1:
            dist.all_reduce(out, op=dist.ReduceOp.SUM)
        return out / float(samples)

    def local_sum(self, x, samples, rank, ws, draw_offset):
        """This rank's part of sum_s softmax: [B, K].  Engines that take a unit window (MCEngine) get the balanced (sample, image)
        split of shard_units; the others whole samples (shard_range)."""
        kw = {} if draw_offset is None else {"draw_offset": draw_offset}
        if getattr(self.engine, "supports_window", False) and not self.engine.regression:
            start, count, first, end = shard_units(samples, x.shape[0], rank, ws)
            if count > 0:
                return self.engine.predict_sum(x, count, sample0=start, window=(first, end), **kw)
        else:
            start, count = shard_range(samples, rank, ws)
            if count > 0:
                return self.engine.predict_sum(x, count, sample0=start, **kw)
        return torch.zeros((x.shape[0], self._n_classes(x)), dtype=torch.float32, device=x.device)

    def predict_async(self, x, samples, then=None, draw_offset=None):
        """Throughput form of predict() for a loop over batches (classification, CUDA): the rank's share of the samples is
        launched on the current stream; the collective, the 1/S scaling and the optional consumer `then(p_bar)` (a metric
        update, a D2H copy) run on a side stream, so the next batch's passes start while this batch's all-reduce is in flight.
        Returns (p_bar, event): p_bar is valid for a stream that has waited on the event (wait_pending() does it for the
        current stream)."""
        if self.engine.regression:
            raise NotImplementedError("predict_async: classification engines (regression heads reduce three sums: use predict)")
        if not x.is_cuda:
            raise RuntimeError("predict_async needs CUDA tensors (streams); predict() is the synchronous form")
        rank, ws = world()
        out = self.local_sum(x, samples, rank, ws, draw_offset)
        cur = torch.cuda.current_stream(x.device)
        side = self.__dict__.get("_side")
        if side is None:
            side = self._side = torch.cuda.Stream(x.device)
        side.wait_stream(cur)
        out.record_stream(side)
        with torch.cuda.stream(side):
            if ws > 1:
                dist.all_reduce(out, op=dist.ReduceOp.SUM)
            p_bar = out / float(samples)
            if then is not None:
                then(p_bar)
            ev = torch.cuda.Event()
            ev.record(side)
        self._last_event = ev
        return p_bar, ev

    def wait_pending(self):
        """Make the current stream wait for every predict_async issued so far."""
        ev = self.__dict__.get("_last_event")
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)

    def _n_classes(self, x):
        """Width of the probability rows, for a rank that owns no sample (more ranks than samples)."""
        k = getattr(self.engine, "n_classes", None)
        if k is None:
            model = self.engine.model
            k = getattr(model, "output_size", None)
            if k is None:
                last = [m for m in model.modules() if hasattr(m, "out_features")]
                k = last[-1].out_features
        return int(k)


def allreduce_prob_sums(psum):
    """The single collective of the sample-sharded eval path."""
    _, ws = world()
    if ws > 1:
        dist.all_reduce(psum, op=dist.ReduceOp.SUM)
    return psum


def reduce_regression(sum_mu, sum_mu2, sum_var, samples):
    """Regression heads under sample sharding: allreduce the three running sums, then
    mean = sum_mu/S ; var = (sum_mu2 - S*mean^2)/(S-1) + sum_var/S  (experiments/utils.py:349-353)."""
    _, ws = world()
    buf = torch.stack([sum_mu, sum_mu2, sum_var])
    if ws > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    mean = buf[0] / samples
    var = (buf[1] - samples * mean * mean) / max(samples - 1, 1) + buf[2] / samples
    return mean, var


def scrub_nan_grads(params):
    """trainer.py:105-107: p.grad[p.grad != p.grad] = 0, per replica, BEFORE the allreduce.  CUDA gradients: one launch for all
    tensors (the pointer list travels in the kernel parameters); anything else: torch, tensor by tensor."""
    import ctypes
    from . import _lib
    cuda = []
    for p in params:
        g = p.grad
        if g is None:
            continue
        if g.is_cuda and g.dtype == torch.float32 and g.is_contiguous():
            cuda.append(g)
        else:
            torch.nan_to_num_(g, nan=0.0, posinf=float("inf"), neginf=float("-inf"))
    if cuda:
        jobs = (_lib.ScrubJob * len(cuda))(*[_lib.ScrubJob(g.data_ptr(), g.numel()) for g in cuda])
        _lib.call("qbn_scrub_nan_multi", jobs, len(cuda), ctypes.c_void_p(torch.cuda.current_stream(cuda[0].device).cuda_stream))


def allreduce_gradients(params, average=True):
    """One flat bucket: flatten -> allreduce(SUM) -> (optionally / world) -> unflatten."""
    _, ws = world()
    grads = [p.grad for p in params if p.grad is not None]
    if ws == 1 or not grads:
        return
    flat = torch._utils._flatten_dense_tensors(grads)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if average:
        flat.div_(ws)
    for g, f in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
        g.copy_(f)


def broadcast_parameters(model, src=0):
    _, ws = world()
    if ws == 1:
        return
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src)


class DPTrainStep:
    """Data-parallel wrapper around the reference's training step (src/trainer.py:87-132): parameters replicated, the batch
    sharded over ranks, ONE flat all-reduce of the gradients per step.

        zero_grad -> output = model(input) -> kl = model.get_kl_divergence() -> criterion(output, target, kl, gamma, n_batches,
        n_points) -> if obj == obj: backward -> NaN-grad scrub (per replica, trainer.py:105-107) -> all-reduce(mean) -> step

    The KL term depends on the parameters only: every rank computes it and scales it with the GLOBAL batch (the criterion's
    kl / (B * n_batches) with 'batch' scaling sees the rank's B, hence `kl_batch_scale = 1 / world`), so it is not reduced twice.
    BatchNorm uses per-rank batch statistics (torch DDP's default); for bit-for-bit parity with a single-GPU run use 1 rank."""

    def __init__(self, model, criterion, optimizer, gamma=0.0, check_nan_loss=True):
        self.model, self.criterion, self.optimizer, self.gamma = model, criterion, optimizer, float(gamma)
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.check_nan_loss = bool(check_nan_loss)
        _, ws = world()
        self.ws = ws
        self.kl_scale = 1.0 / ws
        if ws > 1:
            broadcast_parameters(model)

    def __call__(self, input, target, n_batches, n_points):
        self.optimizer.zero_grad(set_to_none=True)
        output = self.model(input)
        kl = self.model.get_kl_divergence() if hasattr(self.model, "get_kl_divergence") else torch.zeros(1, device=input.device)
        obj, main_obj, kl_term = self.criterion(output, target, kl * self.kl_scale, self.gamma, n_batches, n_points)
        # trainer.py:103 `obj == obj` skips a step whose loss is NaN (one host sync per step, like the reference).  With several
        # ranks the decision must not differ between ranks (a rank that skipped would miss the collective): there every rank
        # always runs the backward; a NaN loss only yields NaN gradients, which the scrub below zeroes before the all-reduce.
        if self.ws > 1 or not self.check_nan_loss or bool(obj == obj):
            obj.backward()
            scrub_nan_grads(self.params)
            allreduce_gradients(self.params, average=True)
            self.optimizer.step()
        return output, obj, main_obj, kl_term


class GraphedTrainStep:
    """The same step as DPTrainStep, captured ONCE in a CUDA graph and replayed: zero_grad -> forward -> KL -> criterion ->
    backward -> NaN-grad scrub -> flat gradient all-reduce -> optimizer.step() (src/trainer.py:87-132).  A B=256 ResNet step is
    ~250 kernel launches of 5-50 us each; eager, the host (autograd + ctypes) is the bottleneck, replayed the GPU never waits.

    * input / target live in static device buffers (`step(input, target)` copies into them: host tensors are fine, pinned is best);
    * fresh noise every replay without re-capturing: the layers' Philox keys are frozen in the graph, the device-side draw
      offset (noise.set_draw_offset) advances by the number of draws of one step before every replay — the draw sequence is the
      one the eager loop would have used;
    * the optimizer must be capturable (torch.optim.Adam(..., capturable=True)); the reference's `obj == obj` host check
      (trainer.py:103) is replaced by the multi-rank rule of DPTrainStep: NaN gradients are zeroed before the all-reduce.
    `warmup` eager steps (real optimisation steps, on the first batch) run before the capture, as torch's capture recipe requires.
    Drop every reference to losses / outputs of earlier eager steps of this model first: a live autograd graph pins the parameters'
    AccumulateGrad nodes to the default stream and invalidates the capture."""

    def __init__(self, model, criterion, optimizer, input, target, n_batches, n_points, gamma=0.0, warmup=3):
        import gc
        from . import noise
        gc.collect()      # an autograd graph of an earlier eager step that is still referenced keeps default-stream AccumulateGrad nodes alive
        for grp in optimizer.param_groups:
            if grp.get("capturable", True) is False:
                raise ValueError("GraphedTrainStep needs a capturable optimizer, e.g. torch.optim.Adam(params, capturable=True)")
        self.eager = DPTrainStep(model, criterion, optimizer, gamma=gamma, check_nan_loss=False)
        self.noise = noise
        self.device = next(model.parameters()).device
        self.input = torch.empty_like(input, device=self.device)
        self.target = torch.empty_like(target, device=self.device)
        self.input.copy_(input)
        self.target.copy_(target)
        self.n_batches, self.n_points = n_batches, n_points
        noise.set_draw_offset(0, self.device)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(max(1, int(warmup))):
                self.eager(self.input, self.target, n_batches, n_points)
        torch.cuda.current_stream(self.device).wait_stream(side)
        self.steps_done = max(1, int(warmup))
        optimizer.zero_grad(set_to_none=True)
        c0 = noise._state["counter"]
        self.graph = torch.cuda.CUDAGraph()
        from . import _lib, config
        _lib.call("qbn_set_pdl", int(config.pdl()))      # programmatic dependent launches of the planar conv kernels inside the graph
        try:
            with torch.cuda.graph(self.graph):
                self.out = self.eager(self.input, self.target, n_batches, n_points)
        finally:
            _lib.call("qbn_set_pdl", 0)
        self.draws_per_step = (noise._state["counter"] - c0) & 0xFFFFFFFF
        self.replays = 0

    def __call__(self, input=None, target=None):
        if input is not None:
            self.input.copy_(input, non_blocking=True)
            self.target.copy_(target, non_blocking=True)
        self.noise.set_draw_offset(self.replays * self.draws_per_step, self.device)
        self.graph.replay()
        self.replays += 1
        self.steps_done += 1
        return self.out
