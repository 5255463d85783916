"""Monte-Carlo prediction of a converted (int8) Bayes-by-backprop model, all samples of a chunk per launch.

The reference evaluates an int8 model like any other: `for i in range(samples): model(x)` (experiments/utils.py:344-355),
every forward re-drawing, re-quantising and re-packing each layer's weights on the CPU.  Here one forward of the *same
model object* carries a chunk of samples: inside `noise.sample_batch` every int8 layer draws its n weight tensors in one
sampler launch and contracts all n samples in one convolution launch; the glue between the layers (`clamp_activation`,
quantised ReLU / residual add / average pool, flatten) is elementwise or per image, so it works unchanged on activations
whose leading dimension is n*B.  Philox streams are keyed by the global sample index: the result does not depend on the
chunk size nor on how samples are sharded over GPUs, and equals a loop of single forwards under `noise.sample_index(s)`
bit for bit (for the usual case B*C*H*W % 64 == 0; see `qbn_i8_add`'s n_vec).

Same interface as `mc.MCEngine` where it matters (`predict`, `predict_sum(x, count, sample0=)`, `regression`), so
`dist.ShardedMCPredictor` shards the samples of an int8 model over GPUs with its single all-reduce."""
import torch

from . import _lib, config, noise, ops
from .stochastic.mcdropout.dropout import BernoulliDropout


def balanced_chunks(total, chunk):
    """Split `total` samples into ceil(total/chunk) chunks whose sizes differ by at most one."""
    if total <= 0:
        return []
    n = -(-total // max(1, chunk))
    base, extra = divmod(total, n)
    return [base + (1 if i < extra else 0) for i in range(n)]


class Int8MCEngine:
    def __init__(self, model, chunk=25, tensor_cores=True):
        layers = [m for m in model.modules() if hasattr(m, "sampled_weights")]
        sites = [m for m in model.modules() if isinstance(m, BernoulliDropout) and m._prob() > 0]
        if not layers and not sites:
            raise ValueError("Int8MCEngine needs a converted model with int8 Bayesian layers (quant_utils.convert) or int8 MC-Dropout "
                             "sites (quant_utils.to_device_int8)")
        self.model, self.chunk, self.regression = model, int(chunk), False
        args = getattr(model, "args", None)
        bits = int(getattr(args, "activation_precision", 8)) if args is not None else 8
        # the model clamps every activation to `bits` right after each layer (clamp_activation, src/utils.py:25-30), so
        # letting the layer clamp to the same width changes nothing — and 7-bit activations qualify for tcgen05 kind::i8
        self.act_bits = bits if tensor_cores else 8

    @torch.no_grad()
    def predict_sum(self, x, samples, sample0=0, seed=None):
        """sum over `samples` MC samples (global indices sample0..) of the model's class probabilities: [B, K] fp32."""
        if not x.is_cuda:
            raise RuntimeError("Int8MCEngine runs on CUDA tensors only (no CPU fallback)")
        if seed is not None:
            noise.manual_seed(seed)
        was_training = self.model.training
        self.model.eval()
        batch, done, total = x.shape[0], 0, None
        try:
            for n in balanced_chunks(int(samples), self.chunk):
                with noise.sample_batch(n, sample0 + done, batch, self.act_bits):
                    probs = self.model(x)                                    # [n*B, K] (or [B, K] if nothing was sampled)
                if probs.shape[0] != n * batch:
                    raise RuntimeError("model output has %d rows for %d samples x batch %d" % (probs.shape[0], n, batch))
                part = ops.mc_mean(probs.reshape(n, batch, -1))
                total = part.mul_(n) if total is None else total.add_(part, alpha=n)
                done += n
        finally:
            self.model.train(was_training)
        return total

    def predict(self, x, samples, sample0=0, seed=None):
        """p-bar = mean over MC samples (experiments/utils.py:355)."""
        return self.predict_sum(x, samples, sample0, seed) / float(samples)


# ------------------------------------------------------------------------------------------------------------------------
class PlanarUnsupported(NotImplementedError):
    """The model does not have the shape the planar int8 path takes; use Int8MCEngine (module-driven, NHWC)."""


class _Conv8:
    __slots__ = ("mod", "src", "dst", "stride", "ksize", "relu", "residual", "add_qp", "add_relu", "sampled", "C", "N", "ref_idx", "name", "dropout")

    def __init__(self, mod, src, dst, ref_idx=None, name=""):
        self.mod, self.src, self.dst, self.ref_idx, self.name = mod, src, dst, ref_idx, name
        self.residual, self.add_qp, self.add_relu = None, None, False
        self.dropout = None          # an int8 MC-Dropout site (dropout.py:31-39) applied to this conv's output, before a residual add
        self.sampled = hasattr(mod, "sampled_weights")
        self.relu = bool(getattr(mod, "RELU", False))
        self.N, self.C = int(mod.out_channels), int(mod.in_channels)
        k, s, p, d = (tuple(v) if isinstance(v, (tuple, list)) else (v, v) for v in (mod.kernel_size, mod.stride, mod.padding, mod.dilation))
        ok = d == (1, 1) and ((k == (3, 3) and p == (1, 1) and s in ((1, 1), (2, 2))) or (k == (1, 1) and p == (0, 0) and s == (2, 2)))
        if not ok or getattr(mod, "groups", 1) != 1:
            raise PlanarUnsupported("conv %s: kernel %s stride %s padding %s dilation %s" % (type(mod).__name__, k, s, p, d))
        self.ksize, self.stride = k[0], s[0]


class Int8PlanarEngine:
    """Monte-Carlo prediction of a converted ResNet-shaped int8 model on the planar zero-copy tcgen05 kernel
    (`qbn_i8_conv_p16_fwd`): quint8 activations live as (q - zero_point) int8 maps in the planar-C16 layout, every conv of
    a chunk of samples is ONE launch whose epilogue does FBGEMM's requantisation, `clamp_activation`, and — for the second
    conv of a BasicBlock — the quantised residual add + ReLU (models_bbb.py:170-183), and the whole S-sample pass is
    replayed from a CUDA graph.  Integers are those of the reference's per-sample CPU loop (experiments/utils.py:344-355
    over conv_q.py:107-125 / linear_q.py:80-94) bit for bit; Philox streams are keyed by the GLOBAL sample index.

    Takes models of the shape  quant -> conv(+ReLU) -> BasicBlocks -> AvgPool2d(whole map) -> Flatten -> linear -> dequant
    whose convs are 3x3/1 (pad 1), 3x3/2 (pad 1) or 1x1/2, with Bayesian int8 layers (conv_q / linear_q) or deterministic
    ones (stochastic.quantized_det: MC-Dropout-free stock nets, SGHMC ensemble members); anything else raises
    PlanarUnsupported and `Int8MCEngine` (module-driven) is the general path.  Same predict / predict_sum interface."""

    def __init__(self, model, chunk=50, use_graph=True):
        self.model, self.chunk, self.use_graph, self.regression = model, int(chunk), bool(use_graph), False
        args = getattr(model, "args", None)
        self.act_bits = int(getattr(args, "activation_precision", 7)) if args is not None else 7
        if self.act_bits > 7:
            raise PlanarUnsupported("activations wider than 7 bits")
        self.steps, self._nreg, self._ref_idx = [], 0, 0
        self.launches = 0
        self.trace = None          # set to {} to keep every step's integer map (quint8 NCHW) of the next eager pass: tests
        self._names = {id(m): n for n, m in model.named_modules()}
        self._compile()

    # ---- compilation
    def _new(self):
        self._nreg += 1
        return self._nreg

    def _conv(self, mod, src):
        """One int8 conv step; Bayesian layers take the next index of the reference's draw order (one eps per layer per forward,
        conv_q.py:113, in call order: stem, stem, shortcut inside a block)."""
        idx = None
        if hasattr(mod, "sampled_weights"):
            idx, self._ref_idx = self._ref_idx, self._ref_idx + 1
        st = _Conv8(mod, src, self._new(), idx, self._names.get(id(mod), ""))
        self.steps.append(st)
        return st

    @staticmethod
    def _is_conv(m):
        return hasattr(m, "in_channels") and hasattr(m, "kernel_size") and hasattr(m, "zero_point") and not isinstance(m, torch.nn.Conv2d)

    @staticmethod
    def _is_linear(m):
        return hasattr(m, "in_features") and hasattr(m, "zero_point") and not isinstance(m, torch.nn.Linear)

    def _compile(self):
        from .quant_utils import DeQuantize, Quantize
        m = self.model
        if not isinstance(getattr(m, "quant", None), Quantize) or not isinstance(getattr(m, "dequant", None), DeQuantize):
            raise PlanarUnsupported("needs a converted model with quant / dequant stubs (quant_utils.convert)")
        cur, tail = 0, []
        flat = []
        for layer in m.layers:
            flat.extend(list(layer) if isinstance(layer, torch.nn.ModuleList) else [layer])
        it = iter(flat)
        for mod in it:
            if isinstance(mod, torch.nn.Identity):
                continue
            if isinstance(mod, BernoulliDropout):
                if mod._prob() > 0:
                    self._attach_dropout(self.steps[-1] if self.steps and self.steps[-1].dst == cur else None, mod)
                continue
            if self._is_conv(mod):
                cur = self._conv(mod, cur).dst
            elif hasattr(mod, "stem") and hasattr(mod, "shortcut") and hasattr(mod, "add"):
                cur = self._block(mod, cur)
            elif isinstance(mod, torch.nn.AvgPool2d):
                tail = [mod] + list(it)
                break
            else:
                raise PlanarUnsupported("module %s in the conv trunk" % type(mod).__name__)
        tail = [t for t in tail if not isinstance(t, torch.nn.Identity)]
        if len(tail) != 3 or type(tail[1]).__name__ != "Flatten" or not self._is_linear(tail[2]):
            raise PlanarUnsupported("tail must be AvgPool2d -> Flatten -> linear")
        self.pool, self.head, self.out_reg = tail[0], tail[2], cur
        self.head_ref_idx = self._ref_idx if hasattr(self.head, "sampled_weights") else None
        self.n_noise = self._ref_idx + (1 if self.head_ref_idx is not None else 0)
        if not self.steps:
            raise PlanarUnsupported("no convolution")
        # a register read by a stride-2 conv is stored phase-split; then every reader must be a stride-2 conv
        self.split = set()
        for st in self.steps:
            if st.stride == 2:
                self.split.add(st.src)
        for st in self.steps:
            if (st.src in self.split and st.stride != 2) or st.residual in self.split or self.out_reg in self.split or 0 in self.split:
                raise PlanarUnsupported("a map feeds both a stride-2 conv and a stride-1 reader")

    @staticmethod
    def _attach_dropout(step, mod):
        """An MC-Dropout site directly behind a conv: one elementwise launch on the conv's planar output (qbn_i8_p16_dropout)."""
        fn = mod.mul_mask
        if step is None or step.dropout is not None:
            raise PlanarUnsupported("an MC-Dropout site that does not directly follow a convolution")
        if not (hasattr(fn, "scale") and hasattr(fn, "zero_point")):
            raise PlanarUnsupported("BernoulliDropout.mul_mask is not converted (quant_utils.convert)")
        step.dropout = mod

    def _block(self, blk, cur):
        def seq(mods, reg):
            last = None
            for mod in mods:
                if isinstance(mod, torch.nn.Identity):
                    continue
                if isinstance(mod, BernoulliDropout):
                    if mod._prob() > 0:
                        self._attach_dropout(last, mod)
                    continue
                if not self._is_conv(mod):
                    raise PlanarUnsupported("module %s inside a BasicBlock" % type(mod).__name__)
                last = self._conv(mod, reg)
                reg = last.dst
            return reg, last
        first = len(self.steps)
        out, last = seq(blk.stem, cur)
        n_stem = len(self.steps) - first
        sc = cur
        if len(blk.shortcut) > 0:
            sc, _ = seq(blk.shortcut, cur)
            moved = self.steps[first + n_stem:]
            del self.steps[first + n_stem:]
            self.steps[first:first] = moved                      # the shortcut runs first: the add rides the second stem conv
        fn = blk.add.add
        if last is None or last.relu or not (hasattr(fn, "scale") and hasattr(fn, "zero_point")):
            raise PlanarUnsupported("BasicBlock without a converted residual add")
        last.residual, last.add_qp = sc, (float(fn.scale), int(fn.zero_point))
        last.add_relu = isinstance(getattr(blk, "end", None), torch.nn.ReLU)
        if not last.add_relu:
            raise PlanarUnsupported("BasicBlock without a final ReLU")
        return out

    # ---- execution
    def _buf(self, key, make):
        cache = self.__dict__.setdefault("_bufs", {})
        if key not in cache:
            cache[key] = make()
        return cache[key]

    def _weights(self, st, n, sample0, C_pad, injected=None):
        mod = st.mod
        if st.sampled:
            if injected is not None:                             # parity tests: the reference's own eps, one tensor per layer per sample
                eps = torch.stack([injected[s][st.ref_idx].float().reshape(-1) for s in range(n)]).contiguous()
                w = ops.i8_sample_weights(mod.weight.reshape(-1), mod.std.reshape(-1), mod._sample_params(), n, eps, 0, 0, 0)
                w = w.reshape((n,) + tuple(mod.weight.shape))
            else:
                w = mod.sampled_weights(n, sample0)              # [n, N, C, R, S] int8, the reference's draw order and arithmetic
            out = self._buf(("w", id(st), n), lambda: torch.empty((n, ops.p16_weight_bytes(C_pad, st.N, st.ksize, st.ksize, st.stride)),
                                                                  dtype=torch.int8, device=w.device))
            self.launches += 2
            return ops.i8_p16_block_weights(w, C_pad, st.stride, out), False
        key = ("wdet", id(st), mod.weight.data_ptr(), mod.weight._version)
        cache = self.__dict__.setdefault("_bufs", {})
        if key not in cache:                                     # fixed weights: blocked once, shared by every sample
            cache[key] = ops.i8_p16_block_weights(mod.weight.reshape((1,) + tuple(mod.weight.shape)), C_pad, st.stride)
        return cache[key], True

    def _chunk_masks(self, n, B, sample0, device):
        """Keep / drop decisions of every MC-Dropout site for the chunk, {id(module): fp32 [n*B, C]}: ONE launch, Philox keyed
        (seed, site id, GLOBAL sample index + draw offset) — the draws BernoulliDropout._forward_int8 makes under noise.sample_batch."""
        sites = [st for st in self.steps if st.dropout is not None]
        if not sites:
            return {}
        from ._lib import MaskJob
        ps = {round(st.dropout._prob(), 9) for st in sites}
        if len(ps) != 1:
            raise PlanarUnsupported("MC-Dropout sites with different probabilities")
        key = ("masks", n, B)
        cache = self.__dict__.setdefault("_bufs", {})
        if key not in cache:
            bufs = {id(st.dropout): torch.empty((n * B, st.N), dtype=torch.float32, device=device) for st in sites}
            jobs = (MaskJob * len(sites))()
            for i, st in enumerate(sites):
                jobs[i] = MaskJob(bufs[id(st.dropout)].data_ptr(), B * st.N, int(st.dropout._qbn_layer_id), 0)
            raw = torch.frombuffer(bytearray(bytes(jobs)), dtype=torch.uint8).to(device)
            cache[key] = (raw, bufs, max(B * st.N for st in sites))
            for st in sites:                               # the multiplier is read once (host value of a frozen parameter)
                st.dropout._mult = float(st.dropout.multiplier.detach().reshape(-1)[0])
        raw, bufs, max_elems = cache[key]
        ops.dropout_masks_multi(raw, len(sites), max_elems, n, 1.0 - next(iter(ps)), noise.seed(), sample0)
        self.launches += 1
        return bufs

    def _run_chunk(self, x, n, sample0, injected=None):
        from .quant_utils import QTensor
        B = x.shape[0]
        bits = self.act_bits
        xq = self.model.quant(x)
        q = torch.clamp(xq.q, 0, (1 << bits) - 1) if xq.bits > bits else xq.q          # clamp_activation (src/utils.py:25-30)
        H, W = q.shape[2], q.shape[3]
        m0 = self._buf(("in", B, q.shape[1], H, W), lambda: ops.P16Map.empty(B, q.shape[1], H + 1, W + 1, 1, x.device))
        regs = {0: ops.P16Map.from_quint8(q, xq.scale, xq.zero_point, bits, out=m0)}
        self.launches += 2
        shared = {0: True}
        masks = self._chunk_masks(n, B, sample0, x.device)
        for si, st in enumerate(self.steps):
            src = regs[st.src]
            if st.stride == 2:
                Ho, Wo = src.Hp - 1, src.Wp - 1                  # the phase maps have the output's geometry
            else:
                Ho, Wo = src.Hp - 1, src.Wp - 1
            split = st.dst in self.split
            if split and (Ho % 2 or Wo % 2):
                raise PlanarUnsupported("phase-split output needs even H, W")
            if st.C != src.C:
                raise RuntimeError("conv expects %d input channels, the map has %d" % (st.C, src.C))
            wb, w_shared = self._weights(st, n, sample0, src.C_pad, injected)
            mod = st.mod
            s_w, z_w = (mod.add_qp if st.sampled else mod.w_qp)
            out = self._buf(("map", si, n, B, Ho, Wo), lambda: (ops.P16Map.empty(n * B, st.N, Ho // 2 + 1, Wo // 2 + 1, 4, x.device) if split
                                                                 else ops.P16Map.empty(n * B, st.N, Ho + 1, Wo + 1, 1, x.device)))
            res = regs[st.residual] if st.residual is not None else None
            if res is not None and shared.get(st.residual, False):
                raise PlanarUnsupported("residual taken from the network input")
            drop = st.dropout
            if drop is None:
                ops.i8_conv_p16_forward(src, wb, n, st.N, st.ksize, st.ksize, st.stride, mod.bias(), s_w, z_w, mod.scale, mod.zero_point, st.relu,
                                        bits, out, residual=res, add_qp=st.add_qp, add_relu=st.add_relu, x_shared=shared[st.src], w_shared=w_shared,
                                        out_phase_split=split)
                self.launches += 1
            else:
                # conv -> (BN, ReLU folded) -> MC-Dropout [-> residual add -> ReLU] (models_mc.py:117-157): the dropout sits between the
                # conv and the add, so conv = requantise only, then ONE elementwise launch: mask multiply (+ add + ReLU).  A conv with
                # fixed weights on an input shared by all samples runs once; its output forks into the samples at the dropout.
                once = shared[st.src] and w_shared
                nc = 1 if once else n
                pre = self._buf(("pre", si, nc, B, Ho, Wo), lambda: ops.P16Map.empty(nc * B, st.N, Ho + 1, Wo + 1, 1, x.device))     # normal layout
                ops.i8_conv_p16_forward(src, wb, nc, st.N, st.ksize, st.ksize, st.stride, mod.bias(), s_w, z_w, mod.scale, mod.zero_point, st.relu,
                                        bits, pre, x_shared=shared[st.src], w_shared=w_shared, out_phase_split=False)
                if res is not None and (res.C * (res.Hp - 1) * (res.Wp - 1) * B) % 64 != 0:
                    raise PlanarUnsupported("fused quantized::add needs maps whose element count is a multiple of 64 (ATen's vector body)")
                ops.i8_p16_dropout(pre, masks[id(drop)], n, B, float(drop.mul_mask.scale), int(drop.mul_mask.zero_point),
                                   float(drop.multiplier.detach().reshape(-1)[0]) if not hasattr(drop, "_mult") else drop._mult, bits, out,
                                   x_shared=once and n > 1, residual=res, add_qp=st.add_qp, add_relu=st.add_relu)
                self.launches += 2
            regs[st.dst] = out
            shared[st.dst] = False
            if self.trace is not None:
                self.trace[st.name] = (out.to_quint8().clone(), out.scale, out.zero_point)
        top = regs[self.out_reg]
        k = self.pool.kernel_size if isinstance(self.pool.kernel_size, int) else self.pool.kernel_size[0]
        if k != top.Hp - 1 or k != top.Wp - 1:
            raise PlanarUnsupported("AvgPool2d(%d) on a %dx%d map: only the global pool is built" % (k, top.Hp - 1, top.Wp - 1))
        pooled = QTensor(ops.i8_p16_avgpool(top, bits), top.scale, top.zero_point, bits)
        if self.trace is not None:
            self.trace["pool"] = (pooled.q.clone(), pooled.scale, pooled.zero_point)
        if injected is not None and self.head_ref_idx is not None:
            ys = []
            for s in range(n):                                   # injected noise is per forward: the head runs sample by sample
                with noise.inject([injected[s][self.head_ref_idx]]):
                    ys.append(self.head(QTensor(pooled.q[s * B:(s + 1) * B], pooled.scale, pooled.zero_point, bits)))
            y = QTensor(torch.cat([t.q.reshape(B, -1) for t in ys]), ys[0].scale, ys[0].zero_point, bits)
        else:
            with noise.sample_batch(n, sample0, B, bits):
                y = self.head(pooled)                            # [n*B, K] quint8: one sampler + one contraction launch
        self.launches += 4
        if self.trace is not None:
            self.trace["head"] = (y.q.clone(), y.scale, y.zero_point)
        return y.dequantize().reshape(n, B, -1)

    supports_window = True

    def _predict_sum_eager(self, x, samples, sample0, injected=None, window=None):
        psum, done = None, 0
        sizes = balanced_chunks(int(samples), self.chunk)
        nB = x.shape[0]
        for ci, n in enumerate(sizes):
            # the call's unit window (dist.shard_units) restricts the first sample of the first chunk and the last of the last chunk
            win = None
            if window is not None:
                win = (window[0] if ci == 0 else 0, window[1] if ci == len(sizes) - 1 else nB)
                if win == (0, nB):
                    win = None
            if win is not None:
                _lib.call("qbn_p4_set_window", win[0], win[1] if win[1] < nB else 0, n)
            try:
                logits = self._run_chunk(x, n, sample0 + done, injected[done:done + n] if injected is not None else None)
            finally:
                if win is not None:
                    _lib.call("qbn_p4_set_window", 0, 0, 0)
            psum = ops.softmax_accumulate(logits.contiguous(), psum, win)
            self.launches += 1
            done += n
        return psum

    @torch.no_grad()
    def predict_sum(self, x, samples, sample0=0, seed=None, injected=None, draw_offset=None, window=None):
        """sum over `samples` MC samples (global indices sample0..) of the class probabilities: [B, K] fp32.
        draw_offset: first global sample index of this batch's draws (fresh noise per batch from the one captured graph).
        injected: per sample, the list of eps tensors of every Bayesian layer in the reference's draw order (parity tests).
        window = (first_img, end_img): the first sample contributes images [first_img, B) only, the last [0, end_img) only."""
        if not x.is_cuda:
            raise RuntimeError("Int8PlanarEngine runs on CUDA tensors only (no CPU fallback)")
        if window is not None:
            window = (int(window[0]), int(window[1]))
            if window == (0, x.shape[0]):
                window = None
            elif not (0 <= window[0] < x.shape[0] and 0 < window[1] <= x.shape[0] and (samples > 1 or window[0] < window[1])):
                raise ValueError("unit window %r outside a batch of %d images" % (window, x.shape[0]))
        if seed is not None:
            noise.manual_seed(seed)
        noise.draw_base(x.device)
        if draw_offset is not None:
            noise.set_draw_offset(draw_offset, x.device)
        x = x.float()
        if not self.use_graph or injected is not None or self.trace is not None:
            return self._predict_sum_eager(x, samples, sample0, injected, window)
        key = (tuple(x.shape), int(samples), int(sample0), noise.seed(), x.device.index, window)
        graphs = self.__dict__.setdefault("_graphs", {})
        ent = graphs.get(key)
        if ent is None:
            static_x = x.clone()
            cur, side = torch.cuda.current_stream(), torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):                        # warm-up: allocates the cached buffers, sets kernel attributes
                self._predict_sum_eager(static_x, samples, sample0, None, window)
            cur.wait_stream(side)
            torch.cuda.synchronize()
            l0 = self.launches
            g = torch.cuda.CUDAGraph()
            _lib.call("qbn_set_pdl", int(config.pdl()))      # programmatic dependent launches between the captured convs
            try:
                with torch.cuda.graph(g):
                    static_out = self._predict_sum_eager(static_x, samples, sample0, None, window)
            finally:
                _lib.call("qbn_set_pdl", 0)
            ent = (g, static_x, static_out, self.launches - l0)
            self.launches = l0
            graphs[key] = ent
            if len(graphs) > 16:
                graphs.pop(next(iter(graphs)))
        g, static_x, static_out, n_launch = ent
        static_x.copy_(x)
        g.replay()
        self.launches += n_launch
        return static_out.clone()

    def predict(self, x, samples, sample0=0, seed=None, injected=None, draw_offset=None):
        return self.predict_sum(x, samples, sample0, seed, injected, draw_offset) / float(samples)


class Int8EnsembleEngine:
    """SGHMC ensemble evaluation (models_sgld.py:216-288): `Network(training_mode=False)` holds `args.samples` independently
    trained members and its forward visits `ensemble[counter]` once per call, so the reference's S-loop
    (experiments/utils.py:344-347) averages one deterministic forward per member.  Here every member is a converted stock
    int8 network (quant_utils.to_device_int8) compiled onto the planar kind::i8 kernel with its weights blocked once; members
    are independent, so `dist.ShardedMCPredictor` shards them over GPUs like Monte-Carlo samples (member index = sample index)
    and all-reduces the probability sums once.  predict_sum(x, count, sample0) = sum over members sample0 .. sample0+count-1."""

    def __init__(self, members, use_graph=True):
        members = list(getattr(members, "ensemble", members))
        if not members:
            raise ValueError("empty ensemble")
        self.members = members
        self.engines = []
        for m in members:
            try:
                self.engines.append(Int8PlanarEngine(m, chunk=1, use_graph=use_graph))
            except PlanarUnsupported:
                self.engines.append(None)                       # module-driven forward of that member (any architecture)
        self.regression, self.model = False, members[0]
        self.n_members = len(members)

    @torch.no_grad()
    def predict_sum(self, x, samples, sample0=0, **_):
        if not x.is_cuda:
            raise RuntimeError("Int8EnsembleEngine runs on CUDA tensors only (no CPU fallback)")
        total = None
        for i in range(int(sample0), int(sample0) + int(samples)):
            k = i % self.n_members                              # models_sgld.py:277-284: the counter wraps around
            eng = self.engines[k]
            if eng is not None:
                part = eng.predict_sum(x, 1)
            else:
                was = self.members[k].training
                self.members[k].eval()
                part = torch.softmax(self.members[k](x.float()), dim=-1)
                self.members[k].train(was)
            total = part.clone() if total is None else total.add_(part)
        return total

    def predict(self, x, samples=None, sample0=0):
        samples = self.n_members if samples is None else samples
        return self.predict_sum(x, samples, sample0) / float(samples)


def make_int8_engine(model, chunk=50, **kw):
    """The fastest engine that takes `model`: the planar tcgen05 path when the model has its shape, else the module-driven one."""
    try:
        return Int8PlanarEngine(model, chunk=chunk, **{k: v for k, v in kw.items() if k in ("use_graph",)})
    except PlanarUnsupported:
        return Int8MCEngine(model, chunk=min(chunk, 25), **{k: v for k, v in kw.items() if k in ("tensor_cores",)})
