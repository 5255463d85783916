"""The Monte-Carlo loop of experiments/utils.py:330-377, re-designed for one B200.

The reference runs S sequential Python-level forwards, each re-sampling every layer's weights
through 4 elementwise kernels, appends S [B,10] outputs to a list, stacks and averages.  Here the
model is compiled once into a short plan of fused steps and ALL samples of a chunk advance through
a layer together:

    sample_weights_blocked_multi   W[s] = mu + sigma*eps_s for EVERY layer of the chunk (Philox, one launch)
    conv_p4_forward (per layer)    y[s] = act(bn(conv(x[s], W[s])) (*mask) + residual[s])   planar-C4, tcgen05, zero-copy im2col
    softmax_accumulate             psum += sum_s softmax(logits[s])                          (no [B,S,10] stack)

Layers the planar kernel does not take (3-channel first layer when it cannot be sample-stacked, LeNet's
1/20-channel convs, linear layers, fp32 mode) run on the gather / fp32 kernels over NHWC tensors.
BatchNorm(eval), bias, ReLU, the residual add, the BasicBlock's 1x1 stride-2 shortcut and MC-Dropout masks are
folded into the conv launches (_plan_p4).  The prepared operands are cached by parameter version and the whole
S-sample pass is replayed from a CUDA graph.

Noise is keyed by the GLOBAL sample index, so any sharding of the S samples over GPUs (dist.py)
reproduces the single-GPU result up to summation order.
"""
import contextlib

import torch
import torch.nn as nn

from . import _lib, config, noise, ops
from ._lib import QBN_MATH_FP32, QBN_MATH_TF32
from .stochastic.bbb.conv import Conv2d as BBBConv2d
from .stochastic.bbb.linear import Linear as BBBLinear


class _ConvStep:
    """conv/linear -> [BN] -> [ReLU (relu_pre)] -> [MC-Dropout of the output] -> [+ residual] -> [ReLU (relu)].
    det: a plain nn.Conv2d / nn.Linear (MC-Dropout networks) = a Bayesian layer with sigma == 0."""
    __slots__ = ("mod", "src", "dst", "bn", "relu", "relu_pre", "dropout", "residual", "ref_idx", "is_linear", "flatten_hw", "det")

    def __init__(self, mod, src, dst, ref_idx, det=False):
        self.mod, self.src, self.dst, self.ref_idx = mod, src, dst, ref_idx
        self.bn, self.relu, self.relu_pre, self.residual = None, False, False, None
        self.dropout = None            # (p, site layer id, ref_idx of its mask draw)
        self.is_linear = isinstance(mod, (BBBLinear, nn.Linear))
        self.flatten_hw = False
        self.det = det


class _PoolStep:
    __slots__ = ("kind", "src", "dst")

    def __init__(self, kind, src, dst):
        self.kind, self.src, self.dst = kind, src, dst


def _is_bbb(m):
    return isinstance(m, (BBBConv2d, BBBLinear))


def _is_det(m):
    return isinstance(m, (nn.Conv2d, nn.Linear)) and not _is_bbb(m)


class MCEngine:
    """Compiled S-sample predictive pass for the reference's BBB model families
    (LinearNetwork / ConvNetwork_LeNet / ConvNetwork_ResNet of models_bbb.py, or their qbn_b200.zoo
    mirrors).  predict() == `_evaluate_with_loader`'s inner loop for one batch."""

    def __init__(self, model, math_mode="tf32", chunk=50, use_graph=True, chunk_max=None, lanes=None, sample_ahead=False):
        self.model = model
        # concurrent chunk streams of a call (1 = off).  Measured on the ResNet, B=256: 2 lanes -1.8 % at 13 samples per call (the
        # per-rank share at 8 GPUs), neutral at 100 — the per-launch cost is CTA start-up work, not idle SMs — so it stays opt-in
        self.lanes = int(lanes) if lanes else 1
        # draw chunk i+1's weights on a second stream under chunk i's convolutions.  Measured on the ResNet, B=256, S=100: 17.315
        # against 17.310 ms — the sampler's blocks find no room beside the resident conv CTAs — so it stays opt-in
        self.sample_ahead = bool(sample_ahead)
        self.use_graph = bool(use_graph)
        self.chunk_max = int(chunk_max) if chunk_max else int(chunk)   # a call's samples are split into ceil(S / chunk_max) balanced chunks
        self.math_mode = {"fp32": QBN_MATH_FP32, "tf32": QBN_MATH_TF32}[math_mode] if isinstance(math_mode, str) else math_mode
        self.chunk = int(chunk)
        self.steps = []
        self.regression = hasattr(model, "mu") and hasattr(model, "log_var")
        self._nreg = 0
        self._ref_idx = 0
        self._compile(model)
        self.n_noise = self._ref_idx
        self.launches = 0
        self._prepared = None

    # ---- compilation ---------------------------------------------------------------------------
    def _new(self):
        self._nreg += 1
        return self._nreg

    def _conv(self, mod, src):
        det = _is_det(mod)
        st = _ConvStep(mod, src, self._new(), None if det else self._ref_idx, det)
        if not det:
            self._ref_idx += 1          # one noise draw per Bayesian layer per forward (conv.py:34-35)
        self.steps.append(st)
        return st

    def _seq(self, mods, cur):
        """Fuse Conv -> [BN] -> [ReLU] runs; returns (last register, last conv step or None)."""
        last = None
        for m in mods:
            if _is_bbb(m) or _is_det(m):
                last = self._conv(m, cur)
                cur = last.dst
            elif type(m).__name__ == "BernoulliDropout":
                if float(m._p) > 0.0:
                    assert last is not None and last.dropout is None, "MC-Dropout must follow a conv/linear (+BN, +ReLU)"
                    last.dropout = (float(m._p), m._qbn_layer_id, self._ref_idx)      # one mask draw per site per forward (dropout.py:21-30)
                    self._ref_idx += 1
                    if last.relu:                  # conv-BN-ReLU-dropout: the ReLU precedes the mask
                        last.relu, last.relu_pre = False, True
            elif isinstance(m, nn.BatchNorm2d):
                assert last is not None and last.bn is None and not last.relu and last.dropout is None, "BatchNorm must follow a conv"
                last.bn = m
            elif isinstance(m, nn.ReLU):
                assert last is not None and not last.relu and last.dropout is None, "ReLU must follow a conv/linear layer"
                last.relu = True
            elif isinstance(m, nn.MaxPool2d):
                assert m.kernel_size in (2, (2, 2)) and m.stride in (2, (2, 2)), "only 2x2/2 max-pool (models_bbb.py:106-108)"
                dst = self._new()
                self.steps.append(_PoolStep("max", cur, dst))
                cur, last = dst, None
            elif isinstance(m, nn.AvgPool2d):
                dst = self._new()
                self.steps.append(_PoolStep("avg", cur, dst))
                cur, last = dst, None
            elif type(m).__name__ in ("Flatten", "Identity", "QuantStub", "DeQuantStub"):
                last = None
            elif isinstance(m, nn.ModuleList):
                for blk in m:
                    cur = self._block(blk, cur)
                last = None
            elif hasattr(m, "stem") and hasattr(m, "shortcut"):
                cur = self._block(m, cur)
                last = None
            else:
                raise NotImplementedError("MCEngine: unsupported module %s" % type(m).__name__)
        return cur, last

    def _block(self, blk, cur):
        """BasicBlock (models_bbb.py:170-183): relu(stem(x) + shortcut(x)); reference noise order is
        stem.0, stem.3, shortcut.0 but the shortcut is EXECUTED first so that the residual add and
        the final ReLU ride the epilogue of the second stem conv."""
        first = len(self.steps)
        ref0 = self._ref_idx
        out, last = self._seq(list(blk.stem), cur)
        n_stem = len(self.steps) - first
        sc = cur
        if len(blk.shortcut) > 0:
            sc, _ = self._seq(list(blk.shortcut), cur)
            sc_steps = self.steps[first + n_stem:]
            del self.steps[first + n_stem:]
            self.steps[first:first] = sc_steps      # run the shortcut first
        assert last is not None and not last.relu
        last.residual = sc
        last.relu = True
        assert self._ref_idx > ref0
        return out

    def _compile(self, model):
        cur = 0  # register 0 = network input
        if self.regression:
            cur, _ = self._seq(list(model.layers), cur)
            self.head_mu = self._conv(model.mu, cur)
            self.head_lv = self._conv(model.log_var, cur)
            self.out_reg = None
        else:
            cur, _ = self._seq(list(model.layers), cur)
            self.out_reg = cur

    # ---- parameter preparation: packed / blocked operands, folded BatchNorm; redone only when a parameter changes ----
    def _param_versions(self):
        ver = []
        for st in self.steps:
            if not isinstance(st, _ConvStep):
                continue
            ts = [st.mod.weight, getattr(st.mod, "std", None), getattr(st.mod, "bias", None)]
            if st.bn is not None:
                ts += [st.bn.weight, st.bn.bias, st.bn.running_mean, st.bn.running_var]
            ver.extend((t.data_ptr(), t._version) for t in ts if t is not None)
        return tuple(ver)

    def _get_prep(self, device):
        """The prepared operands live in persistent tensors, so an unchanged model costs nothing per call and the CUDA
        graph captures only the per-batch work; an optimizer step / load_state_dict bumps the version counters."""
        ver = (self._param_versions(), str(device), self.math_mode)
        cached = self.__dict__.get("_prep_cache")
        if cached is None or cached[0] != ver:
            self.__dict__.pop("_graphs", None)          # graphs hold pointers into the old operands
            self.__dict__.pop("_p4_jobs", None)
            self._prep_cache = (ver, self._prepare(device))
        return self._prep_cache[1]

    def _prepare(self, device):
        prep = {}
        for st in self.steps:
            if not isinstance(st, _ConvStep):
                continue
            m = st.mod
            w, rho = m.weight.detach(), (None if st.det else m.std.detach())
            prep[id(st)] = {"w": w, "rho": rho}
            scale = shift = None
            if st.bn is not None:
                bn = st.bn
                scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).detach().float().contiguous()
                shift = (bn.bias - bn.running_mean * scale).detach().float().contiguous()
                if m.bias is not None:
                    shift = shift + m.bias.detach() * scale
            elif m.bias is not None:
                shift = m.bias.detach().float().contiguous()
            prep[id(st)].update(scale=scale, shift=shift)
        return prep

    def _packed(self, st, prep, x, cmult=4):
        """Pack (mu, sigma) for the geometry this step sees (depends on the input's H, W for the
        flatten->linear case, which runs as an HxW 'valid' convolution over the NHWC map)."""
        e = prep[id(st)]
        key = ("packed", tuple(x.shape[1:]) if st.is_linear else None, cmult)     # only flatten->linear depends on the map size
        if key in e:
            return e[key]
        m = st.mod
        w, rho = e["w"], e["rho"]
        det = rho is None
        if det:
            rho = torch.zeros_like(w)              # placeholder; sigma is forced to exactly 0 below
        flat = None
        if st.is_linear:
            if x.dim() == 4 and (x.shape[2] > 1 or x.shape[3] > 1) and self.math_mode == QBN_MATH_TF32 and x.shape[1] % 4 != 0:
                # flatten -> linear over a map whose channel count is not a multiple of 4 (LeNet: 50 x 7 x 7): run it as a
                # 1x1 layer over the NHWC-flattened vector, K padded to a multiple of 4, so it stays on the tcgen05 path
                C, H, W = x.shape[1], x.shape[2], x.shape[3]
                K = C * H * W
                Kp = (K + 3) // 4 * 4
                wf = w.reshape(w.shape[0], C, H, W).permute(0, 2, 3, 1).reshape(w.shape[0], K)
                rf = rho.reshape(w.shape[0], C, H, W).permute(0, 2, 3, 1).reshape(w.shape[0], K)
                w4 = torch.nn.functional.pad(wf, (0, Kp - K)).reshape(w.shape[0], Kp, 1, 1)
                rho4 = torch.nn.functional.pad(rf, (0, Kp - K), value=-200.0).reshape(w.shape[0], Kp, 1, 1)
                flat = (K, Kp, H * W, (w.shape[0], C, H, W))
                stride, pad, dil = (1, 1), (0, 0), (1, 1)
            elif x.dim() == 4 and (x.shape[2] > 1 or x.shape[3] > 1):
                C, H, W = x.shape[1], x.shape[2], x.shape[3]
                w4, rho4 = w.reshape(w.shape[0], C, H, W), rho.reshape(w.shape[0], C, H, W)
                stride, pad, dil = (1, 1), (0, 0), (1, 1)
            else:
                w4, rho4 = w.reshape(w.shape[0], -1, 1, 1), rho.reshape(w.shape[0], -1, 1, 1)
                stride, pad, dil = (1, 1), (0, 0), (1, 1)
        else:
            w4, rho4 = w, rho
            stride, pad, dil = m.stride, m.padding, m.dilation
        C = w4.shape[1]
        cpad = 0
        if self.math_mode == QBN_MATH_TF32 and C % cmult != 0 and C < cmult:
            cpad = cmult - C % cmult   # 3-channel input layer: pad channels so 16-byte K-chunks stay inside a tap (8 for the planar kernel)
            w4 = torch.nn.functional.pad(w4, (0, 0, 0, 0, 0, cpad))
            rho4 = torch.nn.functional.pad(rho4, (0, 0, 0, 0, 0, cpad), value=-200.0)  # softplus -> 0
        packed = ops.weight_prep(w4.contiguous(), rho4.contiguous(), False, None, want=("mu", "sigma"))
        if cpad or flat:
            packed["sigma"] = torch.where(packed["sigma"] < 1e-30, torch.zeros_like(packed["sigma"]), packed["sigma"])
        if det:
            packed["sigma"] = torch.zeros_like(packed["sigma"])       # W[s] = mu + 0 * eps = mu for every sample
        info = dict(mu=packed["mu"], sigma=packed["sigma"], wshape=tuple(w4.shape), stride=stride, pad=pad, dil=dil, cpad=cpad, flat=flat,
                    orig_shape=tuple(w.shape) if not st.is_linear else (w.shape[0], w4.shape[1] - cpad, w4.shape[2], w4.shape[3]))
        e[key] = info
        return info

    # ---- layout planning: which registers live in the zero-bordered layout -------------------------
    def _s1_eligible(self, st):
        """Stride-1 'same' conv that the zero-copy-im2col kernel (qbn_conv_s1_fwd) takes."""
        if not isinstance(st, _ConvStep) or st.is_linear or self.math_mode != QBN_MATH_TF32:
            return None
        m = st.mod
        R, S_ = m.kernel_size
        if tuple(m.stride) != (1, 1) or tuple(m.dilation) != (1, 1) or R % 2 == 0 or S_ % 2 == 0:
            return None
        if tuple(m.padding) != ((R - 1) // 2, (S_ - 1) // 2) or m.in_channels % 4 != 0 or m.out_channels > 256:
            return None
        if R == 1 and S_ == 1:
            return None
        return ((R - 1) // 2, (S_ - 1) // 2)

    def _plan_layout(self):
        if hasattr(self, "_reg_pad"):
            return self._reg_pad
        producers = {}
        for st in self.steps:
            producers[st.dst] = st
        want = {}
        for st in self.steps:
            pad = self._s1_eligible(st)
            if pad is not None:
                want[st.src] = pad
        changed = True
        while changed:           # residual and output of a fused add share one geometry
            changed = False
            for st in self.steps:
                if isinstance(st, _ConvStep) and st.residual is not None:
                    a, b = st.dst, st.residual
                    for u, v in ((a, b), (b, a)):
                        if u in want and v not in want:
                            want[v] = want[u]
                            changed = True
        def ok(reg):
            pr = producers.get(reg)
            if not isinstance(pr, _ConvStep) or pr.is_linear:
                return False            # network input / pooled maps are not written zero-bordered
            for st in self.steps:       # every consumer must understand the layout
                if isinstance(st, _PoolStep) and st.src == reg and st.kind == "max":
                    return False
            return True
        bad = {r for r in want if not ok(r)}
        changed = True
        while changed:
            changed = False
            for st in self.steps:
                if isinstance(st, _ConvStep) and st.residual is not None:
                    if (st.dst in bad) != (st.residual in bad) and (st.dst in want or st.residual in want):
                        bad.update((st.dst, st.residual))
                        changed = True
        self._reg_pad = {r: p for r, p in want.items() if r not in bad}
        return self._reg_pad

    # ---- planar-C4 planning: which convs run on qbn_conv_p4_fwd and which registers are planar -------
    def _p4_eligible(self, st):
        """('s1', pad) / ('s2', None) if qbn_conv_p4_fwd takes this conv (include/qbn.h, planar-C4 path)."""
        if not isinstance(st, _ConvStep) or st.is_linear or self.math_mode != QBN_MATH_TF32:
            return None
        m = st.mod
        R, S_ = m.kernel_size
        if m.in_channels % 8 != 0 or m.out_channels % 4 != 0 or m.out_channels > 256 or tuple(m.dilation) != (1, 1):
            return None
        if tuple(m.stride) == (1, 1) and R % 2 == 1 and S_ % 2 == 1 and R * S_ > 1 and tuple(m.padding) == ((R - 1) // 2, (S_ - 1) // 2):
            return ("s1", ((R - 1) // 2, (S_ - 1) // 2))
        if tuple(m.stride) == (2, 2) and ((R, S_) == (3, 3) and tuple(m.padding) == (1, 1) or (R, S_) == (1, 1) and tuple(m.padding) == (0, 0)):
            return ("s2", None)
        return None

    def _plan_p4(self):
        """Fixpoint: start from 'every eligible conv runs planar', drop convs / registers until every planar
        register has a producer that can write its layout and only consumers that read it.
        Returns (reg -> ('p4', border) | ('p4s', (1, 1)), set of conv-step ids on the planar kernel)."""
        if hasattr(self, "_p4_plan"):
            return self._p4_plan
        convs = [st for st in self.steps if isinstance(st, _ConvStep)]
        producers = {st.dst: st for st in self.steps}
        on = {id(st): self._p4_eligible(st) for st in convs}
        on = {k: v for k, v in on.items() if v is not None}
        while True:
            need = {}          # reg -> set of layouts demanded by planar convs
            for st in convs:
                e = on.get(id(st))
                if e is None:
                    continue
                need.setdefault(st.src, set()).add(("p4", e[1]) if e[0] == "s1" else ("p4s", (1, 1)))
                out_border = e[1] if e[0] == "s1" else (1, 1)
                if st.residual is not None:
                    need.setdefault(st.residual, set()).add(("p4", out_border))
            layout = {}
            for r, kinds in need.items():
                if len(kinds) == 1:
                    layout[r] = next(iter(kinds))
            # a planar conv's output is planar too: its layout is what its consumers need (default: its own border)
            for st in convs:
                e = on.get(id(st))
                if e is not None and st.dst not in layout and st.dst not in need:
                    layout[st.dst] = ("p4", e[1] if e[0] == "s1" else (1, 1))
            bad_regs = set(r for r, kinds in need.items() if len(kinds) != 1)
            for r, lay in list(layout.items()):
                pr = producers.get(r)
                ok = isinstance(pr, _ConvStep) and not pr.is_linear and self.math_mode == QBN_MATH_TF32
                if ok and id(pr) in on:
                    e = on[id(pr)]
                    own = e[1] if e[0] == "s1" else (1, 1)
                    ok = lay[0] == "p4s" or lay[1] == own          # a planar conv writes its own geometry (or phase-split)
                elif ok:
                    ok = lay[0] == "p4" and pr.mod.out_channels % 4 == 0     # gather kernel: QBN_FLAG_OUT_P4, interior only
                for st in self.steps:                                # every consumer must read this layout
                    if isinstance(st, _PoolStep) and st.src == r:
                        ok = ok and st.kind == "avg" and lay[0] == "p4"
                    elif isinstance(st, _ConvStep):
                        if st.src == r and id(st) not in on:
                            ok = False
                        if st.residual == r and id(st) not in on:
                            ok = False
                if r == getattr(self, "out_reg", None) or (self.regression and r in (self.head_mu.src,)):
                    ok = False
                if not ok:
                    bad_regs.add(r)
            drop = set()
            for st in convs:
                e = on.get(id(st))
                if e is None:
                    continue
                regs_used = [st.src, st.dst] + ([st.residual] if st.residual is not None else [])
                if any(r in bad_regs or r not in layout for r in regs_used):
                    drop.add(id(st))
                elif layout[st.dst][0] == "p4s":
                    H_even = True   # checked at run time (needs even H, W)
            if not drop and not bad_regs:
                break
            if not drop:          # registers went bad but no conv depends on them any more
                for r in bad_regs:
                    layout.pop(r, None)
                break
            for k in drop:
                on.pop(k)
        # first layer (shared input, <= 8 channels): sample-stacked planar launch instead of the gather kernel
        self._p4_first = None
        for st in convs:
            m = st.mod
            if (st.src == 0 and not st.is_linear and id(st) not in on and layout.get(st.dst, (None, None)) == ("p4", (1, 1)) and st.residual is None
                    and m.in_channels <= 8 and tuple(m.kernel_size) == (3, 3) and tuple(m.stride) == (1, 1) and tuple(m.padding) == (1, 1)
                    and tuple(m.dilation) == (1, 1) and m.out_channels % 4 == 0):
                self._p4_first = id(st)
        # BasicBlock downsampling shortcut (1x1 stride 2 -> BN) fused into the second stem conv as extra K blocks
        self._p4_fused, self._p4_skip = {}, set()
        for st in convs:
            if (id(st) not in on or on[id(st)][0] != "s1" or st.residual is None or st.dropout is not None or st.bn is None
                    or layout.get(st.dst, ("p4",))[0] != "p4"):
                continue
            sc = producers.get(st.residual)
            if (isinstance(sc, _ConvStep) and id(sc) in on and on[id(sc)][0] == "s2" and tuple(sc.mod.kernel_size) == (1, 1) and sc.bn is not None
                    and sc.residual is None and sc.dropout is None and not sc.relu and not sc.relu_pre
                    and sum(1 for t in self.steps if isinstance(t, _ConvStep) and (t.src == sc.dst or t.residual == sc.dst)) == 1
                    and not any(isinstance(t, _PoolStep) and t.src == sc.dst for t in self.steps)
                    and ops.p4_shortcut_block_channels(st.mod.in_channels, sc.mod.in_channels) > 0):
                self._p4_fused[id(st)] = sc
                self._p4_skip.add(id(sc))
        self._p4_plan = (layout, set(on))
        return self._p4_plan

    def _p4_weights(self, st, prep, info, stride, cb=0):
        """mu / sigma blocked once per call for the planar kernel (they are shared by all samples).  The blocked
        buffers are persistent (stable pointers for the multi-layer sampler's job table and for CUDA graphs) and
        refreshed in place, so parameter updates between calls are honoured."""
        e = prep[id(st)]
        key = ("p4w", stride, cb)
        if key not in e:
            N, C, R, S_ = info["wshape"]
            store = self.__dict__.setdefault("_p4_wbufs", {})
            if (id(st), cb) not in store:
                nfl = ops.p4_weight_floats(C, N, R, S_, stride) if cb == 0 else (C // cb) * R * S_ * (cb // 4) * ((N + 15) // 16 * 16) * 4
                store[(id(st), cb)] = (torch.empty((1, nfl), dtype=torch.float32, device=info["mu"].device),
                                       torch.empty((1, nfl), dtype=torch.float32, device=info["mu"].device),
                                       torch.ones(((N + 15) // 16 * 16,), dtype=torch.float32, device=info["mu"].device))
            mu_b, sg_b, sc_b = store[(id(st), cb)]
            ops.p4_block_weights(info["mu"], N, C, R * S_, stride, out=mu_b, cb=cb)
            ops.p4_block_weights(info["sigma"], N, C, R * S_, stride, out=sg_b, cb=cb)
            if prep[id(st)]["scale"] is not None:
                sc_b[:N].copy_(prep[id(st)]["scale"])          # persistent copy: the sampler's job table keeps this pointer
            e[key] = (mu_b[0], sg_b[0], sc_b)
        return e[key]

    def _p4_sample_all(self, n, sample0, prep, seed, p4_convs, device, injected=None):
        """ONE sampling launch for every planar conv of the chunk (qbn_sample_weights_blocked_multi).
        injected: per-sample lists of noise tensors in the reference's draw order (parity tests) -> eps pointers.
        A conv with a fused shortcut and the shortcut itself are two jobs filling ONE weight tensor per sample
        ([main blocks][shortcut blocks]), both carrying their BatchNorm scale."""
        from ._lib import P4SampleJob
        first = self._p4_first
        fused, skip = self._p4_fused, self._p4_skip
        fused_of = {id(sc): st for st_id, sc in fused.items() for st in self.steps if id(st) == st_id}     # shortcut -> main step

        def stack_ok(st):
            return id(st) == first

        def stack_groups(st):
            """(first sample, count) groups of the chunk whose stacked channels fit one accumulator tile (256 columns)."""
            gmax = max(1, 256 // st.mod.out_channels)
            ng = (n + gmax - 1) // gmax
            sizes_ = [n // ng + (1 if i < n % ng else 0) for i in range(ng)]
            return [(sum(sizes_[:i]), sizes_[i]) for i in range(ng)]
        steps = [st for st in self.steps if id(st) in p4_convs or (isinstance(st, _ConvStep) and stack_ok(st))]
        tables = self.__dict__.setdefault("_p4_jobs", {})

        def pinfo(st):
            return self._packed(st, prep, torch.empty((0, st.mod.in_channels, 1, 1), device="meta"), 8 if id(st) == first else 4)

        def cb_of(st):
            return ops.p4_shortcut_block_channels(fused_of[id(st)].mod.in_channels, st.mod.in_channels) if id(st) in skip else 0
        for st in steps:                       # refresh the blocked parameters (a no-op unless the parameters changed since the last call)
            self._p4_weights(st, prep, pinfo(st), st.mod.stride[0], cb_of(st))
        if n <= 0:                             # refresh only (two-lane calls block the parameters once, before the fork)
            return None
        # a chunk running beside another one (two lanes), or sampled ahead under its predecessor, owns its weight tensors
        tkey = (n, self.__dict__.get("_slot", 0), self.__dict__.get("_wslot", 0))
        if tkey not in tables or injected is not None:
            n_jobs_total = sum(len(stack_groups(st)) if stack_ok(st) else 1 for st in steps)
            jobs = (P4SampleJob * n_jobs_total)()
            ji = 0
            wbufs, max_fl = {}, 0
            keep_eps = []
            sizes = {id(st): self._p4_weights(st, prep, pinfo(st), st.mod.stride[0], cb_of(st))[0].numel() for st in steps}
            for st in steps:                   # weight tensors: fused pairs share one [n, main + shortcut] tensor
                if id(st) in skip:
                    continue
                if stack_ok(st):   # one blocked tensor per group of samples, stacked along N (padding rows stay zero)
                    N, C, R, S_ = pinfo(st)["wshape"]
                    wbufs[id(st)] = [(s0_, cnt, torch.zeros((1, ops.p4_weight_floats(C, cnt * N, R, S_, 1)), dtype=torch.float32, device=device))
                                     for s0_, cnt in stack_groups(st)]
                else:
                    extra = sizes[id(fused[id(st)])] if id(st) in fused else 0
                    wbufs[id(st)] = torch.empty((n, sizes[id(st)] + extra), dtype=torch.float32, device=device)
            for i, st in enumerate(steps):
                info = pinfo(st)
                N, C, R, S_ = info["wshape"]
                mu_b, sg_b, sc_b = self._p4_weights(st, prep, info, st.mod.stride[0], cb_of(st))
                max_fl = max(max_fl, mu_b.numel())
                eps_ptr = None
                if injected is not None and not st.det:
                    es = []
                    for s_ in range(n):
                        e_ = injected[s_][st.ref_idx].reshape(info["orig_shape"]).float()
                        if info["cpad"]:
                            e_ = torch.nn.functional.pad(e_, (0, 0, 0, 0, 0, info["cpad"]))
                        es.append(ops.pack_ohwi(e_))
                    keep_eps.append(torch.stack(es).contiguous())
                    eps_ptr = keep_eps[-1].data_ptr()
                scale_ptr, stride4, w_ptr = None, 0, None
                if id(st) in skip:             # shortcut half of a fused pair: after the main blocks of every sample
                    main = fused_of[id(st)]
                    w = wbufs[id(main)]
                    w_ptr, stride4, scale_ptr = w.data_ptr() + sizes[id(main)] * 4, w.shape[1] // 4, sc_b.data_ptr()
                elif id(st) in fused:
                    w = wbufs[id(st)]
                    w_ptr, stride4, scale_ptr = w.data_ptr(), w.shape[1] // 4, sc_b.data_ptr()
                elif stack_ok(st):
                    for s0_, cnt, wg in wbufs[id(st)]:
                        jobs[ji] = P4SampleJob(mu_b.data_ptr(), sg_b.data_ptr(), eps_ptr, wg.data_ptr(), N, C, R * S_, st.mod.stride[0],
                                               getattr(st.mod, "_qbn_layer_id", 0), cnt, None, 0, 0, s0_, 0)
                        ji += 1
                    continue
                else:
                    w_ptr = wbufs[id(st)].data_ptr()
                jobs[ji] = P4SampleJob(mu_b.data_ptr(), sg_b.data_ptr(), eps_ptr, w_ptr, N, C, R * S_, st.mod.stride[0],
                                       getattr(st.mod, "_qbn_layer_id", 0), 0, scale_ptr, cb_of(st), stride4, 0, 0)
                ji += 1
            raw = torch.frombuffer(bytearray(bytes(jobs)), dtype=torch.uint8).to(device)
            entry = (raw, n_jobs_total, max_fl, wbufs)
            if injected is None:
                tables[tkey] = entry
            else:
                self._p4_keep = (keep_eps, entry)          # keep the eps tensors alive until the chunk has run
        else:
            entry = tables[tkey]
        raw, n_jobs, max_fl, wbufs = entry
        ops.sample_weights_blocked_multi(raw, n_jobs, max_fl, n, seed, sample0, True)
        self.launches += 1
        return wbufs

    def _chunk_masks(self, n, B, sample0, seed, injected, device):
        """MC-Dropout masks of every site for the chunk: {id(step): ([n*B, C] mask, multiplier)}.
        Philox: one launch for all sites, keyed (seed, site id, GLOBAL sample index); injected: the reference's draws."""
        sites = [st for st in self.steps if isinstance(st, _ConvStep) and st.dropout is not None]
        if not sites:
            return {}
        from ._lib import MaskJob
        out = {}
        if injected is not None:
            for st in sites:
                p_, _, ridx = st.dropout
                m = torch.stack([injected[s_][ridx].float().reshape(B, -1) for s_ in range(n)]).reshape(n * B, -1).contiguous().to(device)
                out[id(st)] = (m, float(torch.ones(1) / (1.0 - torch.ones(1) * p_)))
            return out
        cache = self.__dict__.setdefault("_mask_jobs", {})
        key = (n, B, self.__dict__.get("_slot", 0))
        if key not in cache:
            by_p = {}
            for st in sites:
                C = st.mod.out_channels if not st.is_linear else st.mod.out_features
                buf = torch.empty((n * B, C), dtype=torch.float32, device=device)
                by_p.setdefault(st.dropout[0], []).append((st, buf, C))
            groups = []
            for p_, lst in by_p.items():
                jobs = (MaskJob * len(lst))()
                for i, (st, buf, C) in enumerate(lst):
                    jobs[i] = MaskJob(buf.data_ptr(), B * C, st.dropout[1], 0)
                raw = torch.frombuffer(bytearray(bytes(jobs)), dtype=torch.uint8).to(device)
                groups.append((p_, raw, len(lst), max(B * c for _, _, c in lst), lst))
            cache[key] = groups
        for p_, raw, n_jobs, max_elems, lst in cache[key]:
            keep = float(torch.ones(1) - torch.ones(1) * p_)           # dropout.py:21: bernoulli_(1. - self.p), fp32
            ops.dropout_masks_multi(raw, n_jobs, max_elems, n, keep, seed, sample0)
            self.launches += 1
            mult = float(torch.ones(1) / (1.0 - torch.ones(1) * p_))    # dropout.py:10
            for st, buf, C in lst:
                out[id(st)] = (buf, mult)
        return out

    def _p4_buffer(self, key, n_img, C, Hp, Wp, border, phases, device, zero):
        cache = self.__dict__.setdefault("_bufs", {})
        k = (key, n_img, C, Hp, Wp, phases, self.__dict__.get("_slot", 0))
        if k not in cache:
            cache[k] = ops.P4Map.empty(n_img, C, Hp, Wp, border, phases, device, zero=zero)
        return cache[k]

    def _buffer(self, key, shape, device, zero):
        """Output buffers are cached per (step, chunk size): stable pointers, no allocator churn, and the
        zero border of a zero-bordered map is written only once."""
        cache = self.__dict__.setdefault("_bufs", {})
        k = (key, tuple(shape), self.__dict__.get("_slot", 0))
        if k not in cache:
            buf = torch.empty(shape, dtype=torch.float32, device=device, memory_format=ops.CL if len(shape) == 4 else torch.contiguous_format)
            cache[k] = buf.zero_() if zero else buf
        return cache[k]

    # ---- execution -------------------------------------------------------------------------------
    def _prestage(self, x, prep):
        """The planar copy of the shared input batch (first layer on the planar kernel), made once per call."""
        p4_layout, p4_convs = self._plan_p4()
        first = getattr(self, "_p4_first", None)
        if not p4_convs or first is None or x.dim() != 4:
            return
        st = next(s_ for s_ in self.steps if id(s_) == first)
        info = self._packed(st, prep, torch.empty((0, st.mod.in_channels, 1, 1), device="meta"), 8)
        self._call["x_p4"] = ops.p4_stage_input(x, info["wshape"][1], (1, 1))
        self.launches += 1
        self._p4_sample_all(0, 0, prep, noise.seed(), p4_convs, x.device)       # blocked mu / sigma of every planar layer

    def _run_chunk(self, x, n, sample0, prep, injected, presampled=None):
        """Advance samples [sample0, sample0+n) through every step.  Returns the head output(s)."""
        reg_pad = self._plan_layout()
        p4_layout, p4_convs = self._plan_p4()
        if p4_convs:
            reg_pad = {r: v for r, v in reg_pad.items() if r not in p4_layout}
        self._p4_presampled = None
        if p4_convs:
            # presampled: this chunk's weights were drawn ahead, on the sampler stream, under the previous chunk's convolutions
            self._p4_presampled = presampled if presampled is not None else self._p4_sample_all(n, sample0, prep, noise.seed(), p4_convs, x.device, injected)
        masks = self._chunk_masks(n, x.shape[0], sample0, noise.seed(), injected, x.device)
        pending = {}         # register -> (mask, mult): MC-Dropout applied in the operand load of its consumers (gather / fp32 kernels)
        regs = {0: x}
        shared = {0: True}
        ready = {0: False}   # register holds TF32-exact values (written by a TF32 epilogue with OUT_ROUND_TF32)
        B = x.shape[0]
        seed = noise.seed()
        for si, st in enumerate(self.steps):
            src = regs[st.src]
            spad = reg_pad.get(st.src, (0, 0)) if st.src in reg_pad else (0, 0)
            if isinstance(st, _PoolStep):
                if isinstance(src, ops.P4Map):
                    bh, bw = src.border
                    regs[st.dst] = ops.avgpool_p4(src, float((src.Hp - bh) * (src.Wp - bw))).reshape(src.n_img, src.C, 1, 1)
                elif st.kind == "max":
                    regs[st.dst] = ops.maxpool2x2(src)
                else:
                    interior = (src.shape[2] - 2 * spad[0]) * (src.shape[3] - 2 * spad[1])
                    regs[st.dst] = ops.avgpool_all(src, float(interior)).reshape(src.shape[0], src.shape[1], 1, 1)
                shared[st.dst] = shared[st.src]
                ready[st.dst] = ready[st.src] and st.kind == "max"   # max of TF32-exact values is TF32-exact
                if st.src in pending:      # x*m*c with m in {0,1}, c > 0 commutes with max- and average-pooling
                    pending[st.dst] = pending[st.src]
                self.launches += 1
                continue
            if self._p4_presampled is not None and id(st) in self._p4_presampled and id(st) == self._p4_first:
                # first layer on the planar kernel: the shared input is staged once, the chunk's samples are stacked along N
                info = self._packed(st, prep, torch.empty((0, st.mod.in_channels, 1, 1), device="meta"), 8)
                N, C, R, S_ = info["wshape"]
                if "x_p4" not in self._call:
                    # one launch: channel padding, RNA rounding to TF32 (the tensor core would truncate), planar layout with its borders
                    self._call["x_p4"] = ops.p4_stage_input(src, C, (1, 1))
                    self.launches += 1
                xm = self._call["x_p4"]
                outp = self._p4_buffer(("p4first", si), n * xm.n_img, N, xm.Hp, xm.Wp, (1, 1), 1, src.device, zero=False)
                e = prep[id(st)]
                mk, mult = masks.get(id(st), (None, 1.0))
                for s0_, cnt, wg in self._p4_presampled[id(st)]:      # one launch per group of <= 256 / N stacked samples
                    ops.conv_p4_forward(xm, wg, cnt, N, R, S_, 1, e["scale"], e["shift"], None, st.relu,
                                        ops.QBN_FLAG_OUT_ROUND_TF32 | ops.QBN_FLAG_X_SHARED_STACKED | (ops.QBN_FLAG_RELU_PRE if st.relu_pre else 0),
                                        False, outp.images(s0_ * xm.n_img, cnt * xm.n_img), False,
                                        mk[s0_ * xm.n_img:] if mk is not None else None, mult)
                    self.launches += 1
                regs[st.dst] = outp
                shared[st.dst] = False
                ready[st.dst] = True
                continue
            if id(st) in self._p4_skip and self._p4_presampled is not None:
                continue                       # 1x1 stride-2 shortcut: accumulated inside the block's second stem conv
            if id(st) in p4_convs:
                assert st.src not in pending, "planar conv fed by a register with a pending dropout mask (planner bug)"
                regs[st.dst] = self._run_p4_conv(st, si, src, regs, n, sample0, prep, injected, p4_layout, seed, masks)
                self.launches += 1 if self._p4_presampled is not None else 2
                shared[st.dst] = False
                ready[st.dst] = True
                continue
            # geometry of this conv on the UNPADDED map
            if src.dim() == 2:
                src = src.reshape(src.shape[0], src.shape[1], 1, 1)
            unp = src
            if spad != (0, 0):
                H0, W0 = src.shape[2] - 2 * spad[0], src.shape[3] - 2 * spad[1]
                unp = torch.empty((0, src.shape[1], H0, W0), device="meta")          # shape carrier only
            info = self._packed(st, prep, unp)
            if info["cpad"]:
                src = torch.nn.functional.pad(src, (0, 0, 0, 0, 0, info["cpad"])).contiguous(memory_format=ops.CL)
            if info["flat"]:
                K_, Kp_, HW_, _ = info["flat"]
                assert spad == (0, 0)
                flat_x = src.permute(0, 2, 3, 1).reshape(src.shape[0], K_)          # NHWC memory order: a view
                src = torch.nn.functional.pad(flat_x, (0, Kp_ - K_)).reshape(src.shape[0], Kp_, 1, 1)
            N, C, R, S_ = info["wshape"]
            nb = src.shape[0] if shared[st.src] else src.shape[0] // n
            eps = None
            if injected is not None and not st.det:
                es = []
                for s in range(n):
                    if info["flat"]:
                        K_, Kp_, HW_, shp = info["flat"]
                        e = ops.pack_ohwi(injected[s][st.ref_idx].reshape(shp).float()).reshape(shp[0], K_)
                        es.append(torch.nn.functional.pad(e, (0, Kp_ - K_)).reshape(-1))
                        continue
                    e = injected[s][st.ref_idx].reshape(info["orig_shape"]).float()
                    if info["cpad"]:
                        e = torch.nn.functional.pad(e, (0, 0, 0, 0, 0, info["cpad"]))
                    es.append(ops.pack_ohwi(e))
                eps = torch.stack(es).contiguous()
            e = prep[id(st)]
            # tcgen05 eligibility: 16-byte K chunks; one accumulator tile holds <= 256 channels, except that linear (1x1 on a 1x1 map)
            # layers of any width are split over N by the library
            lin_geom = R == 1 and S_ == 1 and src.shape[2] == 1 and src.shape[3] == 1
            elig = config.tf32_eligible(C, N, False) or (C % 4 == 0 and lin_geom and st.residual is None)
            mode = self.math_mode if elig else QBN_MATH_FP32
            tf32 = mode == QBN_MATH_TF32
            w = ops.sample_weights(info["mu"], info["sigma"], n, eps, seed, getattr(st.mod, "_qbn_layer_id", 0), sample0, round_tf32=tf32)
            res = regs[st.residual] if st.residual is not None else None
            # MC-Dropout on the gather / fp32 kernels: the mask of the producing site rides the operand load of its consumers
            in_mask, in_mult = pending.get(st.src, (None, 1.0))
            if in_mask is not None and info["flat"]:        # mask[b, c] broadcast over the flattened (h, w, c) vector
                K_, Kp_, HW_, _ = info["flat"]
                in_mask = torch.nn.functional.pad(in_mask.repeat(1, HW_), (0, Kp_ - K_)).contiguous()
            relu_eff = st.relu or st.relu_pre
            if st.dropout is not None and (st.residual is not None or st.dst in p4_layout):
                raise NotImplementedError("MC-Dropout before a residual add needs the planar kernel (epilogue mask)")
            if res is not None and shared.get(st.residual, False):
                res = res.repeat(n, 1, 1, 1).contiguous(memory_format=ops.CL)   # only if a block reads the raw input
            dpad = reg_pad.get(st.dst, (0, 0))
            flags = ops.QBN_FLAG_OUT_ROUND_TF32 if tf32 else 0
            if st.dst in p4_layout:
                # gather kernel writing the planar-C4 layout directly (first layer: shared input, stacked samples)
                border = p4_layout[st.dst][1]
                d = ops.make_desc(nb, src.shape[2], src.shape[3], C, N, R, S_, info["stride"], info["pad"], info["dil"])
                d.out_pad_h, d.out_pad_w = border
                outp = self._p4_buffer(("v1p4", si), n * nb, N, d.Ho + border[0], d.Wo + border[1], border, 1, src.device, zero=True)
                if tf32 and ready[st.src] and not info["cpad"]:
                    flags |= ops.QBN_FLAG_A_TF32_READY
                ops.conv_forward(src, w, d, n, shared[st.src], False, e["scale"], e["shift"], None, relu_eff, in_mask, in_mult, mode, outp.buf,
                                 flags | ops.QBN_FLAG_OUT_P4)
                assert res is None
                self.launches += 2
                regs[st.dst] = outp
                shared[st.dst] = False
                ready[st.dst] = True
                continue
            s1 = self._s1_eligible(st)
            if s1 is not None and tf32 and spad == s1 and dpad == s1 and ready[st.src] and not shared[st.src] and not info["cpad"] and in_mask is None:
                out = self._buffer(("s1", si, n), (src.shape[0], N, src.shape[2], src.shape[3]), src.device, zero=False)
                ops.conv_s1_forward(src, w, n, N, R, S_, e["scale"], e["shift"], res, relu_eff, flags, False, out)
            else:
                pad_eff = (info["pad"][0] - spad[0], info["pad"][1] - spad[1])      # reading the interior of a bordered map
                d = ops.make_desc(nb, src.shape[2], src.shape[3], C, N, R, S_, info["stride"], pad_eff, info["dil"])
                if dpad != (0, 0):
                    if not tf32:
                        raise RuntimeError("zero-bordered output needs the tcgen05 path")
                    d.out_pad_h, d.out_pad_w = dpad
                    out = self._buffer(("v1", si, n), (n * nb, N, d.Ho + 2 * dpad[0], d.Wo + 2 * dpad[1]), src.device, zero=True)
                else:
                    out = self._buffer(("v1", si, n), (n * nb, N, d.Ho, d.Wo), src.device, zero=False)
                if tf32 and ready[st.src] and not info["cpad"]:
                    flags |= ops.QBN_FLAG_A_TF32_READY
                ops.conv_forward(src, w, d, n, shared[st.src], False, e["scale"], e["shift"], res, relu_eff, in_mask, in_mult, mode, out, flags)
            self.launches += 2
            regs[st.dst] = out
            shared[st.dst] = False
            ready[st.dst] = tf32
            if st.dropout is not None:
                pending[st.dst] = masks[id(st)]
        if self.out_reg in pending or (self.regression and (self.head_mu.src in pending)):
            raise NotImplementedError("MC-Dropout on the network output")
        if self.regression:
            return regs[self.head_mu.dst].reshape(n, B), regs[self.head_lv.dst].reshape(n, B)
        logits = regs[self.out_reg]
        return logits.reshape(n, B, -1)

    def _run_p4_conv(self, st, si, src, regs, n, sample0, prep, injected, p4_layout, seed, masks):
        """One BBB conv on the planar-C4 kernel: blocked sampling launch + qbn_conv_p4_fwd."""
        assert isinstance(src, ops.P4Map), "planar conv fed by a non-planar register (planner bug)"
        m = st.mod
        stride = m.stride[0]
        if src.phases == 4:
            H0, W0 = 2 * (src.Hp - 1), 2 * (src.Wp - 1)
        else:
            H0, W0 = src.Hp - src.border[0], src.Wp - src.border[1]
        info = self._packed(st, prep, torch.empty((0, src.C, H0, W0), device="meta"))
        N, C, R, S_ = info["wshape"]
        mu_b, sg_b, _ = self._p4_weights(st, prep, info, stride)
        if self._p4_presampled is not None and id(st) in self._p4_fused:
            # out = relu(conv3x3(y) * s2 + conv1x1/2(x_block) * ssc + (b2 + bsc)): both branches in one accumulator
            sc = self._p4_fused[id(st)]
            e, esc = prep[id(st)], prep[id(sc)]
            out = self._p4_buffer(("p4", si), src.n_img, N, src.Hp, src.Wp, src.border, 1, src.buf.device, zero=False)
            assert p4_layout[st.dst][0] == "p4"
            if "shift_fused" not in e:
                e["shift_fused"] = (e["shift"] + esc["shift"]).contiguous()
            ops.conv_p4_shortcut_forward(src, self._p4_presampled[id(st)], regs[sc.src], n, N, R, S_, None, e["shift_fused"], st.relu,
                                         ops.QBN_FLAG_OUT_ROUND_TF32, out)
            return out
        if self._p4_presampled is not None:
            w = self._p4_presampled[id(st)]
        else:
            eps = None
            if injected is not None:
                eps = torch.stack([ops.pack_ohwi(injected[s][st.ref_idx].reshape(info["orig_shape"]).float()) for s in range(n)]).contiguous()
            w = ops.sample_weights_blocked(mu_b, sg_b, N, C, R * S_, n, eps, seed, m._qbn_layer_id, sample0, True, None, stride)
        e = prep[id(st)]
        lay = p4_layout[st.dst]
        split = lay[0] == "p4s"
        if stride == 2:
            Hp_o, Wp_o, border = src.Hp, src.Wp, (1, 1)
        else:
            Hp_o, Wp_o, border = src.Hp, src.Wp, src.border
        if split:
            Ho, Wo = Hp_o - border[0], Wp_o - border[1]
            if Ho % 2 or Wo % 2:
                raise RuntimeError("phase-split output needs even H, W")
            out = self._p4_buffer(("p4", si), src.n_img, N, Ho // 2 + 1, Wo // 2 + 1, (1, 1), 4, src.buf.device, zero=True)
        else:
            out = self._p4_buffer(("p4", si), src.n_img, N, Hp_o, Wp_o, border, 1, src.buf.device, zero=False)
        res = regs[st.residual] if st.residual is not None else None
        mk, mult = masks.get(id(st), (None, 1.0))
        ops.conv_p4_forward(src, w, n, N, R, S_, stride, e["scale"], e["shift"], res, st.relu,
                            ops.QBN_FLAG_OUT_ROUND_TF32 | (ops.QBN_FLAG_RELU_PRE if st.relu_pre else 0), False, out, split, mk, mult)
        return out

    @torch.no_grad()
    def predict_sum(self, x, samples, sample0=0, injected=None, draw_offset=None, window=None):
        """Sum over samples [sample0, sample0+samples) of softmax(logits) -> [B,K] (classification) or
        (sum mu, sum mu^2, sum var) building blocks (regression: returns stacked [S,B] mu and var).

        window = (first_img, end_img) (classification; dist.shard_units): the first sample contributes images [first_img, B)
        only and the last images [0, end_img) only — the planar conv launches skip the tiles of the other rows.

        With use_graph (classification, Philox noise) the ~45 launches per chunk of a call are captured once per
        (input shape, sample range, seed) into a CUDA graph and replayed: the step is launch-bound otherwise."""
        if not x.is_cuda:
            raise RuntimeError("MCEngine needs CUDA tensors (no CPU fallback)")
        if draw_offset is not None:     # fresh noise for this batch: global sample indices draw_offset + sample0 + s (noise.set_draw_offset)
            noise.set_draw_offset(draw_offset, x.device)
        if window is not None:
            window = (int(window[0]), int(window[1]))
            if window == (0, x.shape[0]):
                window = None
            elif self.regression or not (0 <= window[0] < x.shape[0] and 0 < window[1] <= x.shape[0] and (samples > 1 or window[0] < window[1])):
                raise ValueError("unit window %r outside a batch of %d images (classification only)" % (window, x.shape[0]))
        if self.use_graph and injected is None and not self.regression:
            return self._predict_sum_graph(x, samples, sample0, window)
        return self._predict_sum_eager(x, samples, sample0, injected, window)

    supports_window = True

    def _predict_sum_graph(self, x, samples, sample0, window=None):
        self._get_prep(x.device)        # a parameter update since the capture invalidates the graphs (they replay prepared operands)
        noise.draw_base(x.device)       # the captured samplers read the per-batch draw offset from this device scalar
        key = (tuple(x.shape), x.dtype, int(samples), int(sample0), noise.seed(), x.device.index, window)
        graphs = self.__dict__.setdefault("_graphs", {})
        ent = graphs.get(key)
        if ent is None:
            static_x = x.clone()
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):                      # warm-up: allocates the cached buffers, sets kernel attributes
                self._predict_sum_eager(static_x, samples, sample0, None, window)
                if self.lanes > 1:                             # ... and once more on all lanes (their buffers, their job tables)
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

    def _predict_sum_eager(self, x, samples, sample0=0, injected=None, window=None):
        x = ops.nhwc(x.float()) if x.dim() == 4 else x.float().contiguous().reshape(x.shape[0], -1, 1, 1)
        prep = self._get_prep(x.device)
        self._call = {}                 # per-call (input-dependent) cache, e.g. the planar copy of the input batch
        self._slot = 0                  # buffer set of the lane a chunk runs on (reset: an exception may have left another lane's)
        psum = None
        mus, lvs = [], []
        done = 0
        # balanced chunks: ceil(S / chunk) passes of near-equal size.  Every launch has ~14 us of ramp / tail, so fewer, larger
        # chunks win (measured at S=100: chunk 10 -> 19.8 ms, 20 -> 18.6, 50 -> 18.2, 100 -> 18.6); the first layer's
        # sample-stacked launch is split into groups of <= 256 / N samples inside a chunk
        n_chunks = (samples + self.chunk_max - 1) // self.chunk_max
        # two lanes: the chunks alternate between two streams (a fork / join inside a captured graph), each with its own activation
        # and weight buffers, so the ramp of one chunk's launch fills the tail of the other's — every launch is persistent and
        # occupies the whole GPU, the block scheduler hands the SMs over as the CTAs of the finishing launch retire
        lanes = self.lanes if (injected is None and not self.regression and samples >= 2 * self.lanes and x.is_cuda) else 1
        # the first call for an input shape after a parameter change packs / blocks the layers' operands inside the chunk that first
        # needs them: that call runs on one lane, so no other lane can read an operand that is still being written
        lanes_ready = prep.setdefault("_lanes_ready", set())
        if lanes > 1 and tuple(x.shape) not in lanes_ready:
            lanes = 1
        if lanes > 1:
            n_chunks = (n_chunks + lanes - 1) // lanes * lanes
        sizes = [samples // n_chunks + (1 if i < samples % n_chunks else 0) for i in range(n_chunks)]
        nB = x.shape[0]
        cur = torch.cuda.current_stream(x.device) if lanes > 1 else None
        streams, psums = [None], [None] * lanes
        if lanes > 1:
            self._prestage(x, prep)            # the shared planar input is staged once, before the fork
            streams = self.__dict__.setdefault("_lane_streams", {}).setdefault(x.device.index, [torch.cuda.Stream(x.device) for _ in range(lanes)])
            for st_ in streams:
                st_.wait_stream(cur)
        # sample ahead: chunk i+1's weights are drawn on a second stream (a second branch of the captured graph) while chunk i's
        # convolutions run — double-buffered weight tensors; only the first chunk's sampling stays on the critical path
        ahead = (self.sample_ahead and lanes == 1 and injected is None and len(sizes) > 1 and x.is_cuda and bool(self._plan_p4()[1]))
        pending = None
        if ahead:
            cur = torch.cuda.current_stream(x.device)
            sampler = self.__dict__.setdefault("_sampler_streams", {}).setdefault(x.device.index, torch.cuda.Stream(x.device))
        for ci, n in enumerate(sizes):
            inj = injected[done:done + n] if injected is not None else None
            presampled = None
            if ahead:
                p4_convs = self._plan_p4()[1]
                if pending is None:
                    self._wslot = ci % 2
                    presampled = self._p4_sample_all(n, sample0 + done, prep, noise.seed(), p4_convs, x.device)
                else:
                    presampled, ev = pending
                    cur.wait_event(ev)
                pending = None
                if ci + 1 < len(sizes):
                    sampler.wait_stream(cur)           # chunk i-1 (the last reader of these weight tensors) is already in the stream
                    with torch.cuda.stream(sampler):
                        self._wslot = (ci + 1) % 2
                        nxt = self._p4_sample_all(sizes[ci + 1], sample0 + done + n, prep, noise.seed(), p4_convs, x.device)
                        ev = torch.cuda.Event()
                        ev.record(sampler)
                    pending = (nxt, ev)
                self._wslot = 0
            # the call's unit window restricts the first sample of the first chunk and the last sample of the last chunk
            win = None
            if window is not None:
                win = (window[0] if ci == 0 else 0, window[1] if ci == len(sizes) - 1 else nB)
                if win == (0, nB):
                    win = None
            lane = ci % lanes
            self._slot = lane
            with torch.cuda.stream(streams[lane]) if lanes > 1 else contextlib.nullcontext():
                if win is not None:
                    _lib.call("qbn_p4_set_window", win[0], win[1] if win[1] < nB else 0, n)
                try:
                    out = self._run_chunk(x, n, sample0 + done, prep, inj, presampled)
                finally:
                    if win is not None:
                        _lib.call("qbn_p4_set_window", 0, 0, 0)
                if self.regression:
                    mus.append(out[0].clone())      # the step buffers are reused by the next chunk
                    lvs.append(out[1].clone())
                else:
                    psums[lane] = ops.softmax_accumulate(out.contiguous(), psums[lane], win)
                    self.launches += 1
            done += n
        self._slot = 0
        lanes_ready.add(tuple(x.shape))
        psum = psums[0]
        if lanes > 1:
            for st_ in streams:
                cur.wait_stream(st_)
            for extra in psums[1:]:
                if extra is not None:
                    extra.record_stream(cur)
                    psum.record_stream(cur)
                    psum = psum + extra
                    self.launches += 1
        if self.regression:
            return torch.cat(mus), torch.cat(lvs).exp()
        return psum

    @torch.no_grad()
    def predict(self, x, samples, sample0=0, injected=None, draw_offset=None):
        """Classification: mean_s softmax (experiments/utils.py:355).  Regression: (mean, var) of
        experiments/utils.py:349-353.  draw_offset: first global sample index of this batch's draws (a data loader passes
        batch_index * S so that every batch gets fresh noise from the one captured graph)."""
        out = self.predict_sum(x, samples, sample0, injected, draw_offset)
        if self.regression:
            mu, var = out
            self.launches += 1
            return ops.reg_mc_reduce(mu, var)
        return out / float(samples)
