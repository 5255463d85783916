#!/usr/bin/env python
"""bench.py — headline benchmark of the stochastic-layer hot path.

Metric (BASELINE.json): ResNet-18 (24/48/96/192) Bayes-by-backprop MC-sampled images/sec at S=100
on synthetic CIFAR-shape data (B=256 x 3x32x32, random-init trained-like weights).

  python bench.py --gpus N --steps K --warmup W          our arm (one process per GPU under torchrun)
  python bench.py --impl reference ...                   the reference's own CPU path (oracle/_ref, else the oracle port), rank 0 only

One "step" = one pass of the hot path over one batch: 256 images x 100 MC samples -> p-bar -> metrics.
N > 1: the 100 samples are sharded over ranks (global Philox sample index), one NCCL allreduce of
the [256,10] probability sums per step; total work per step is fixed => "scaling": "strong".
Every step draws FRESH noise (the reference redraws per batch, experiments/utils.py:342-347): the per-batch draw offset lives in
a device scalar, so the one captured CUDA graph serves every step; `replay_same_noise` reports the old regime beside it.

Secondary legs in the same JSON line (each can be switched off):
  "train"  config 4: LRT training step of the same network, B=256 per GPU, data-parallel (one flat NCCL gradient all-reduce)
  "int8"   config 5: the A7/W8 converted network on the planar kind::i8 kernel, S=100 sharded like the headline
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B, S, K_CLASSES = 256, 100, 10
FLOP_PER_SAMPLE_IMAGE = 1.5704e8          # SURVEY.md §8d, one contraction per layer
FLOP_TRAIN_PER_IMAGE = 9.396e8            # SURVEY.md §8d, LRT forward (2 contractions) + backward (4)
ACT_BYTES_PER_SAMPLE_IMAGE = 1.929e6      # fp32 activations in+out of the 21 stochastic layers
ACT_BYTES_PER_SAMPLE_IMAGE_U8 = 0.482e6   # the same with quint8 activations
# roofline.traffic: dram__bytes_read.sum + dram__bytes_write.sum per conv launch from ONE `ncu --set full` capture of THIS
# benchmark configuration, kept as a small summary under profiles/ (scripts/ncu_traffic.py writes it from the .ncu-rep).
# Reported only when the summary was captured at the chunking this run uses; otherwise null (never rescaled).
TRAFFIC_SUMMARY = os.path.join(ROOT, "profiles", "r02_p4_dram_traffic.json")
WORKLOAD = "ResNet-18(24/48/96/192) BBB eval, B=256 x 3x32x32, S=100 MC samples"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def _cublas_tf32():
    """cuBLAS TF32 / INT8 GEMM throughput measured on a B200 of this pool (scripts/measure_peaks.py), beside the derived figure."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_measured_peaks_tf32_int8.json")) as f:
            p = json.load(f)
        return {"tf32_tflops": p["tf32_tflops"], "tf32_tflops_sustained": p["tf32_tflops_sustained"], "int8_tops": p["int8_tops"],
                "source": "torch.matmul 8192^3 (scripts/measure_peaks.py, profiles/r02_measured_peaks_tf32_int8.json)"}
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_traffic(samples, chunk, n_launches):
    """(bytes per conv launch, provenance) from the committed ncu summary if it matches this run's chunking, else (None, why)."""
    try:
        with open(TRAFFIC_SUMMARY) as f:
            t = json.load(f)
        if int(t["samples_per_rank"]) == int(samples) and int(t["chunk"]) == int(chunk) and int(t["conv_launches"]) == int(n_launches):
            return float(t["dram_bytes_per_launch"]), t["source"]
        return None, "no ncu capture for samples=%d chunk=%d (summary holds samples=%s chunk=%s)" % (samples, chunk, t.get("samples_per_rank"), t.get("chunk"))
    except Exception as e:
        return None, "no ncu summary (%s)" % type(e).__name__


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own modules (oracle/_ref: verbatim copy of /root/reference made by oracle/build_ref.py) when
# present and intact, else the oracle port (the same torch-CPU operators, restated).  Test / baseline infrastructure only.
# ------------------------------------------------------------------------------------------------------------------
def _reference_model():
    """The unmodified reference's `conv_resnet_bbb` (models_bbb.py:191-259) with the benchmark's seeded parameters, or None."""
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("_qbn_build_ref", os.path.join(ROOT, "oracle", "build_ref.py"))
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
        from oracle import ref_harness
        if not ref_harness.reference_available():
            return None
        if ref_harness.REFERENCE_ROOT.endswith("_ref") and not ref.verify():
            return None
        ref_harness.import_reference()
        import oracle.qbn_oracle as O
        from src.models.stochastic.bbb.models_bbb import ConvNetwork_ResNet
        args = ref_harness.Args(sigma_prior=0.05, model="conv_resnet_bbb", task="classification", samples=S)
        net = ConvNetwork_ResNet([1, 3, 32, 32], K_CLASSES, False, args)
        P = O.ResNetBBBParams(seed=1)
        sd = net.state_dict()
        for name, (mu, rho) in P.convs.items():
            sd[name + ".weight"], sd[name + ".std"] = mu, rho
        for name, (w, b, rm, rv, _) in P.bns.items():
            sd[name + ".weight"], sd[name + ".bias"], sd[name + ".running_mean"], sd[name + ".running_var"] = w, b, rm, rv
        sd["layers.9.weight"], sd["layers.9.std"] = P.fc
        net.load_state_dict(sd)
        return net.eval()
    except Exception as e:                                   # the port below always exists
        sys.stderr.write("reference modules unavailable (%s: %s); timing the oracle port\n" % (type(e).__name__, e))
        return None


def cpu_rate(n_threads, budget_s=12.0, max_samples=40, device="cpu"):
    """images/s at S=100 of the reference's MC loop (experiments/utils.py:342-355) on a bounded sample: S' sequential eval
    forwards of the B=256 batch, scaled to S.  Returns (rate, S', seconds per forward, kind)."""
    import torch
    torch.set_num_threads(n_threads)
    x = torch.randn(B, 3, 32, 32, generator=torch.Generator().manual_seed(2))
    net = _reference_model()
    if net is not None:
        kind = "reference"
        if device != "cpu":
            net, x = net.to(device), x.to(device)
        fwd = lambda: net(x)  # noqa: E731
    else:
        kind = "port"
        import oracle.qbn_oracle as O
        P = O.ResNetBBBParams(seed=1)
        eps_fn = lambda name, shape: torch.empty(shape).normal_()  # noqa: E731
        fwd = lambda: O.resnet_bbb_eval_forward(P, x, eps_fn)  # noqa: E731
    sync = (lambda: torch.cuda.synchronize()) if device != "cpu" else (lambda: None)
    with torch.no_grad():
        fwd()
        sync()
        t0 = time.perf_counter()
        n = 0
        while n < max_samples and (time.perf_counter() - t0 < budget_s or n < 3):
            fwd()
            n += 1
        sync()
        dt = time.perf_counter() - t0
    per_sample = dt / n
    return B / (per_sample * S), n, per_sample, kind


def run_reference(args):
    """--impl reference: rank 0 times the reference's CPU path with all host threads; other ranks exit 0."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    vals, sp, kind = [], None, "port"
    cpu_rate(cores, budget_s=2.0, max_samples=3)
    for _ in range(max(1, args.steps)):
        v, n, per, kind = cpu_rate(cores, budget_s=max(2.0, 60.0 / max(1, args.steps)), max_samples=5)
        vals.append(v)
        sp = (n, per)
    value = statistics.mean(vals)
    what = ("the UNMODIFIED reference modules (oracle/_ref: conv_resnet_bbb, model.eval(), the loop of experiments/utils.py:342-347)"
            if kind == "reference" else "oracle port = the torch-CPU operators the reference calls")
    line = {
        "impl": "reference", "metric": "resnet18_bbb_mc_images_per_sec_S100", "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * B / value, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": B, "samples": S},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": torch.get_num_threads(), "kind": kind,
                         "sample": "each step = %d sequential eval forwards of the B=256 batch (%.3f s each) scaled x%d/%d to S=100; %s"
                                   % (sp[0], sp[1], S, sp[0], what)},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
def train_leg(torch, dist, dev, rank, world, steps, math):
    """Config 4: one LRT training step (forward, KL, ELBO, backward, NaN scrub, flat gradient all-reduce, Adam) of the same
    network, B=256 per GPU (weak scaling), through dist.DPTrainStep (= src/trainer.py:87-132 made data-parallel)."""
    from qbn_b200 import config, losses, noise, synthetic, zoo
    from qbn_b200 import dist as qdist
    config.set_math_mode(math)
    model = zoo.resnet_from_params(synthetic.ResNetBBBParams(seed=1)).to(dev).train()
    noise.manual_seed(1234 + rank)                       # independent epsilon substreams per replica
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-3, capturable=True, fused=True)
    args = zoo.Args(loss_multiplier=1.0)
    crit = losses.LOSS_FACTORY["classification"](args, "batch")
    step = qdist.DPTrainStep(model, crit, opt, gamma=0.01, check_nan_loss=False)
    g = torch.Generator().manual_seed(5 + rank)
    x = torch.randn(B, 3, 32, 32, generator=g).to(dev)
    t = torch.randint(0, K_CLASSES, (B,), generator=g).to(dev)

    def timed(fn):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps, out
    for _ in range(3):
        step(x, t, 176, 45000)
    ms_eager, out = timed(lambda: step(x, t, 176, 45000))
    ms_step, mode = ms_eager, "eager (one Python-driven step per iteration)"
    loss_eager = float(out[1].detach())
    out = None        # a live loss keeps the eager autograd graph (and its default-stream AccumulateGrad nodes) alive: capture would fail
    graph_note, gstep = None, None
    try:       # the whole step replayed from one CUDA graph (dist.GraphedTrainStep): fresh noise per replay via the device-side draw offset
        gstep = qdist.GraphedTrainStep(model, crit, opt, x, t, 176, 45000, gamma=0.01, warmup=3)
        for _ in range(3):
            gstep()
        ms_graph, out = timed(lambda: gstep())
        ms_step, mode = ms_graph, "one CUDA graph per step (dist.GraphedTrainStep), fresh noise every replay"
    except Exception as e:      # noqa: BLE001 - reported, never hidden
        graph_note = "graph capture failed (%s: %s); eager number reported" % (type(e).__name__, str(e)[:200])
    peak = _peaks()["bf16_tflops_sustained"] / 2.0
    tf = FLOP_TRAIN_PER_IMAGE * B * world / (ms_step * 1e-3) / 1e12
    loss = float(out[1].detach()) if out is not None else loss_eager
    del model, opt, step, gstep
    return {"metric": "resnet18_bbb_lrt_train_images_per_sec", "value": B * world / (ms_step * 1e-3), "unit": "images/s", "ms_per_step": ms_step,
            "steps": steps, "scaling": "weak", "dtype": math, "loss": loss, "mode": mode, "ms_per_step_eager": ms_eager, "graph_note": graph_note,
            "config": {"workload": "ResNet-18(24/48/96/192) BBB LRT training step, B=256 per GPU, Adam lr 1e-3, gamma .01, n_batches 176",
                       "parallelism": "dp%d, one flat NCCL gradient all-reduce (12.6 MB)" % world},
            "roofline": {"bound": "tensor", "achieved": tf / world, "peak": peak, "unit": "TFLOP/s", "frac": tf / world / peak,
                         "note": "per GPU: 9.396e8 flop per image (SURVEY 8d) / step time; peak = 1/2 x sustained bf16 (TF32 not in MEASURED_PEAKS.json)"}}


def int8_leg(torch, dist, dev, rank, world, steps, chunk):
    """Config 5: the A7/W8 converted network (lifecycle on the device: prepare_model -> calibrate -> convert) through the planar
    kind::i8 engine, S=100 sharded over ranks like the headline, fresh noise every step."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_int8", os.path.join(ROOT, "scripts", "bench_int8.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    from qbn_b200 import dist as qdist
    from qbn_b200 import noise
    from qbn_b200.mc_int8 import make_int8_engine
    net, x, _ = mod.build_model(B)
    noise.manual_seed(20261017)
    eng = make_int8_engine(net, chunk=chunk)
    if getattr(eng, "supports_window", False):            # balanced (sample, image) units, like the headline
        start, count, wf, we = qdist.shard_units(S, B, rank, world)
        kw = {"window": (wf, we)}
        count_eff = ((rank + 1) * S * B // world - rank * S * B // world) / B
    else:
        start, count = qdist.shard_range(S, rank, world)
        kw, count_eff = {}, count

    predictor = qdist.ShardedMCPredictor(eng) if kw else None

    def one(k):
        if predictor is not None:      # the all-reduce of step k on the predictor's side stream, under step k+1 (like the headline)
            return predictor.predict_async(x, S, draw_offset=k * S)[0]         # p-bar, valid after wait_pending()
        psum = eng.predict_sum(x, count, sample0=start, draw_offset=k * S, **kw)
        qdist.allreduce_prob_sums(psum)
        return psum / S
    for k in range(3):
        one(k)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        p = one(3 + k)
    if predictor is not None:
        predictor.wait_pending()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms) / steps
    hbm = ACT_BYTES_PER_SAMPLE_IMAGE_U8 * B * count_eff / (ms_step * 1e-3) / 1e9
    pk = _peaks()
    return {"metric": "resnet18_bbb_int8_mc_images_per_sec_S100", "value": B / (ms_step * 1e-3), "unit": "images/s", "ms_per_step": ms_step,
            "steps": steps, "scaling": "strong", "dtype": "u8 x s8 -> s32 (A7/W8), fp32 requantisation", "engine": type(eng).__name__,
            "row_sum_err": float((p.sum(-1) - 1).abs().max()),
            "config": {"workload": "ResNet-18(24/48/96/192) BBB int8 A7/W8 eval, B=256, S=100 MC samples", "chunk": chunk,
                       "parallelism": "mc-sample sharding x%d" % world},
            "roofline": {"bound": "hbm", "achieved": hbm, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": hbm / pk["hbm_gbs"],
                         "note": "whole step on 0.482 MB per sample-image (SURVEY 8d, quint8 activations); the epilogue is instruction-issue bound, "
                                 "see DESIGN.md"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--chunk", type=int, default=int(os.environ.get("QBN_CHUNK", "50")))
    ap.add_argument("--chunk-max", type=int, default=int(os.environ.get("QBN_CHUNK_MAX", "0")), help="largest chunk (0: same as --chunk)")
    ap.add_argument("--math", default="tf32", choices=["tf32", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the config-4 training leg")
    ap.add_argument("--no-int8", action="store_true", help="skip the config-5 int8 leg")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the reference-modules-on-this-GPU context leg")
    ap.add_argument("--pdl", type=int, default=int(os.environ.get("QBN_PDL", "0")), help="programmatic dependent launch inside the captured graphs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank == 0:
        ge.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    if rank != 0:
        ge.build()
    from qbn_b200 import dist as qdist
    from qbn_b200 import config as qconfig
    from qbn_b200 import mc, metrics, noise, synthetic, zoo   # the GPU arm never imports oracle/ (only the cpu_baseline leg does)
    qconfig.set_pdl(bool(args.pdl))

    dev = torch.device("cuda", local_rank)
    P = synthetic.ResNetBBBParams(seed=1)
    model = zoo.resnet_from_params(P).to(dev).eval()
    noise.manual_seed(20261017)
    engine = mc.MCEngine(model, math_mode=args.math, chunk=args.chunk, chunk_max=args.chunk_max)
    x_host = torch.randn(B, 3, 32, 32, generator=torch.Generator().manual_seed(2)).pin_memory()
    t_host = torch.randint(0, K_CLASSES, (B,), generator=torch.Generator().manual_seed(3)).pin_memory()
    x_dev, t_dev = x_host.to(dev), t_host.to(dev)
    # balanced (sample, image) unit split: S=100 over 8 ranks = 12.5 samples each (dist.shard_units), whole samples when N | S
    start, count, win_first, win_end = qdist.shard_units(S, B, rank, world)
    units = (rank + 1) * S * B // world - rank * S * B // world
    count_eff = units / B                       # samples' worth of work of this rank
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2
    metric = metrics.ClassificationMetric(K_CLASSES, device=dev)
    batch_no = [0]

    def next_offset(fresh=True):
        """First global sample index of this batch's draws: batch k uses k*S .. k*S+S-1 (one device-scalar write, no re-capture)."""
        if fresh:
            batch_no[0] += 1
        return batch_no[0] * S

    predictor = qdist.ShardedMCPredictor(engine)

    def step_resident(fresh=True):
        # the rank's samples on the launch stream; all-reduce + metric update on the predictor's side stream (they overlap the next step)
        return predictor.predict_async(x_dev, S, then=lambda p_bar: metric.update(p_bar, t_dev), draw_offset=next_offset(fresh))[0]

    # e2e: a double-buffered loader loop over HOST batches through the public API (ShardedMCPredictor.predict_async): step k's H2D
    # copy (pinned -> device, copy stream) runs under step k-1's passes, the all-reduce, the metric update and the D2H copies of the
    # probabilities and the metric state run on the predictor's side stream, and the host waits for step k-1's result while step k
    # is in flight.  Every step's input crosses PCIe and every step's result reaches the host inside the timed region.
    copy_stream = torch.cuda.Stream(dev)
    dev_in = [(torch.empty_like(x_dev), torch.empty_like(t_dev)) for _ in range(2)]
    host_out = [(torch.empty((B, K_CLASSES), dtype=torch.float32).pin_memory(), torch.empty_like(metric.state, device="cpu").pin_memory())
                for _ in range(2)]

    def e2e_loop(steps):
        done = []
        for k in range(steps):
            xd, td = dev_in[k % 2]
            with torch.cuda.stream(copy_stream):
                if k >= 2:
                    copy_stream.wait_event(done[k - 2])          # the buffers' previous consumer (step k-2) has finished
                xd.copy_(x_host, non_blocking=True)
                td.copy_(t_host, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(copy_stream)
            torch.cuda.current_stream().wait_event(ready)
            ph, sh = host_out[k % 2]

            def consume(p_bar, td=td, ph=ph, sh=sh):
                metric.update(p_bar, td)
                ph.copy_(p_bar, non_blocking=True)
                sh.copy_(metric.state, non_blocking=True)
            done.append(predictor.predict_async(xd, S, then=consume, draw_offset=next_offset())[1])
            if k >= 1:
                done[k - 1].synchronize()                        # the host has step k-1's probabilities and metric state
        done[-1].synchronize()
        return host_out[(steps - 1) % 2]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, isolated=False):
        """Device time of `steps` steps, max over ranks.  Default: ONE event bracket around the K steps (barrier + synchronize on both
        sides), steps queued back to back as a loop over a data loader issues them; every step streams > 1 GB of activations per
        rank, so nothing of the previous step survives in the 126 MB L2.  isolated=True: every step on an idle GPU with the L2
        flushed first and a barrier + synchronize around it (a latency figure: it contains the host's launch path)."""
        ts = []
        if isolated:
            for _ in range(steps):
                flush.fill_(1.0)
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                predictor.wait_pending()
                e1.record()
                barrier()
                ts.append(e0.elapsed_time(e1))
        else:
            flush.fill_(1.0)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            predictor.wait_pending()
            e1.record()
            barrier()
            ts.append(e0.elapsed_time(e1))
        tot = torch.tensor([sum(ts)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        return float(tot.item())

    for _ in range(max(3, args.warmup)):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- timed: K steps, each bracketed by CUDA events on the launching stream; L2 flushed between steps; fresh noise every step
    engine.launches = 0
    total_ms = timed(step_resident, args.steps)
    launches = engine.launches + 3 * args.steps  # + draw-offset fill, metric kernel (+ allreduce)
    n_graphs = len(engine.__dict__.get("_graphs", {}))
    # the old regime for comparison: every step replays the SAME S draws
    n_replay = max(3, args.steps // 2)
    replay_ms = timed(lambda: step_resident(False), n_replay) / n_replay
    isolated_ms = timed(step_resident, n_replay, isolated=True) / n_replay
    # ---- e2e: host buffers, H2D + D2H inside the timed region (wall clock around the double-buffered loop, max over ranks)
    e2e_loop(2)
    barrier()
    t0 = time.perf_counter()
    e2e_loop(args.steps)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    # the same resident loop once more: the GPU runs into its power cap within the first second of load, so the legs measured later
    # (e2e, roofline) see a ~2 % slower device than the first timed loop; reported as `value_after_e2e`
    resident_again_ms = timed(step_resident, n_replay) / n_replay
    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline of the dominant kernel: per-launch CUDA events on the launch stream
    roof = None
    if rank == 0:
        peaks = _peaks()
        evs = []
        orig = mc.ops.conv_forward

        def timed_conv(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = orig(*a, **k)
            e1.record()
            mode = a[12] if len(a) > 12 else k.get("math_mode", 0)
            d, n = a[2], a[3]
            flops = 2.0 * n * d.B * d.Ho * d.Wo * d.N * d.R * d.S * d.C
            evs.append((e0, e1, flops, mode))
            return out
        orig_p4 = mc.ops.conv_p4_forward

        def timed_p4(x, w, n, N, R, S_, stride=1, *a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = orig_p4(x, w, n, N, R, S_, stride, *a, **k)
            e1.record()
            bh, bw = ((R - 1) // 2, (S_ - 1) // 2) if stride == 1 else (1, 1)
            H, W = x.Hp - bh, x.Wp - bw                             # output pixels (the map geometry is the output's)
            evs.append((e0, e1, 2.0 * x.n_img * H * W * N * R * S_ * x.C, 2))
            return out
        orig_p4sc = mc.ops.conv_p4_shortcut_forward

        def timed_p4sc(x, w, x2, n, N, R, S_, *a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = orig_p4sc(x, w, x2, n, N, R, S_, *a, **k)
            e1.record()
            H, W = x.Hp - (R - 1) // 2, x.Wp - (S_ - 1) // 2
            evs.append((e0, e1, 2.0 * x.n_img * H * W * N * (R * S_ * x.C + x2.C), 2))     # 3x3 conv + the fused 1x1 stride-2 shortcut
            return out
        mc.ops.conv_forward, mc.ops.conv_p4_forward, mc.ops.conv_p4_shortcut_forward = timed_conv, timed_p4, timed_p4sc
        flush.fill_(1.0)
        torch.cuda.synchronize()
        engine.use_graph = False                      # per-launch events need the eager launch sequence (the timed loop replays a CUDA graph)
        engine.predict_sum(x_dev, count, sample0=start, window=(win_first, win_end))
        engine.use_graph = True
        torch.cuda.synchronize()
        mc.ops.conv_forward, mc.ops.conv_p4_forward, mc.ops.conv_p4_shortcut_forward = orig, orig_p4, orig_p4sc
        um = [(a.elapsed_time(b), f) for a, b, f, m in evs if m >= 1]
        n_p4 = sum(1 for e in evs if e[3] == 2)
        if um:
            t_ms = sum(t for t, _ in um)
            fl = sum(f for _, f in um) * (count_eff / count)      # the unit window skips the tiles of the other rank's images
            ach = fl / (t_ms * 1e-3) / 1e12
            peak = peaks["bf16_tflops_sustained"] / 2.0
            alg_bytes = ACT_BYTES_PER_SAMPLE_IMAGE * B * count_eff
            # with fp32 activations the eval path's arithmetic intensity (81 flop/B) is below the TF32 ridge
            # (~105 flop/B at the measured peaks), so HBM is the binding roofline; the tensor-pipe view is reported beside it.
            hbm = alg_bytes / (t_ms * 1e-3) / 1e9
            traffic, traffic_source = measured_traffic(count if count_eff == count else -1, args.chunk_max or args.chunk, len(um))
            roof = {"kernel": "umma_conv_p4_kernel (%d launches) + umma_conv_kernel<EVAL> (%d) — tcgen05 kind::tf32" % (n_p4, len(um) - n_p4),
                    "bound": "hbm", "achieved": hbm, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm / peaks["hbm_gbs"],
                    "traffic": traffic,
                    "launches": len(um), "avg_launch_ms": t_ms / len(um),
                    "peak_source": "hbm_gbs of %s" % peaks["source"],
                    "traffic_source": traffic_source,
                    "algorithmic_bytes_per_launch": alg_bytes / len(um),
                    "algorithmic_bytes_note": "1.929 MB per sample-image (SURVEY 8d: activations in + out of the 21 stochastic layers, fp32) x the "
                                              "sample-images of the step / conv launches",
                    "tensor_view": {"achieved_tflops": ach, "peak_tflops": peak, "frac": ach / peak,
                                    "peak_source": "1/2 x sustained bf16 of %s (TF32 peak not in MEASURED_PEAKS.json)" % peaks["source"],
                                    "cublas_tf32_8192": _cublas_tf32()},
                    "share_of_step": t_ms / (total_ms / args.steps)}
    m = metric.compute() if rank == 0 else None
    n_state = metric.state.numel()
    # ---- secondary legs (all ranks take part: they contain collectives).  The eval engine's buffers are released first.
    del engine
    torch.cuda.empty_cache()
    train = int8 = None
    if not args.no_train:
        try:
            train = train_leg(torch, dist, dev, rank, world, max(3, args.steps), args.math)
        except Exception as e:                       # a secondary leg must not take the headline down
            train = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
        torch.cuda.empty_cache()
    if not args.no_int8:
        try:
            int8 = int8_leg(torch, dist, dev, rank, world, max(3, args.steps), args.chunk)
        except Exception as e:
            int8 = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
        torch.cuda.empty_cache()
    if rank == 0:
        value = B * args.steps / (total_ms * 1e-3)
        e2e_v = B * args.steps / e2e_s
        line = {
            "metric": "resnet18_bbb_mc_images_per_sec_S100", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "tf32" if args.math == "tf32" else "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": B, "samples": S,
                       "parallelism": "mc-sample sharding x%d (balanced (sample, image) units: %.1f samples per rank)" % (world, S / world), "chunk": args.chunk,
                       "noise": "fresh draws every step (device-side draw offset, %d captured graph%s for all steps)" % (n_graphs, "" if n_graphs == 1 else "s"),
                       "l2": "inputs larger than L2: every step streams > 1 GB of activations per rank (S/N x 25 MB per 24-channel layer) through the "
                             "126 MB L2; flushed once before the timed region",
                       "timing": "K steps in one CUDA-event bracket on the launch stream (barrier + synchronize on both sides, max over ranks); "
                                 "the all-reduce and the metric update of step k run on a side stream under step k+1"},
            "sample_images_per_sec": value * S,
            "e2e": {"value": e2e_v, "unit": "images/s", "h2d_bytes_per_step": x_host.numel() * 4 + t_host.numel() * 8,
                    "d2h_bytes_per_step": B * K_CLASSES * 4 + n_state * 4,
                    "mode": "double-buffered loop over pinned host batches through ShardedMCPredictor.predict_async: H2D of step k under step "
                            "k-1, D2H of the probabilities and the metric state on the side stream, the host reads step k-1's result while "
                            "step k runs; wall clock around the K steps"},
            "value_after_e2e": {"value": B / (resident_again_ms * 1e-3), "unit": "images/s", "ms_per_step": resident_again_ms,
                                "note": "the timed resident loop repeated after the e2e leg (device under sw_power_cap by then): the e2e figure is to "
                                        "be read against this one"},
            "isolated_step": {"ms_per_step": isolated_ms, "value": B / (isolated_ms * 1e-3), "unit": "images/s",
                              "note": "every step alone on an idle GPU: L2 flushed, barrier + synchronize around each step (the round-1/early "
                                      "round-2 way of timing; contains the host launch path and the exposed all-reduce)"},
            "replay_same_noise": {"value": B / (replay_ms * 1e-3), "unit": "images/s", "ms_per_step": replay_ms,
                                  "note": "every step replays the same S draws (round-1 regime); fresh/replay step time = %.3f" % ((total_ms / args.steps) / replay_ms)},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof,
            "tensor_bound_frac_whole_step": (FLOP_PER_SAMPLE_IMAGE * B * S * args.steps / (total_ms * 1e-3) / 1e12) / (world * _peaks()["bf16_tflops_sustained"] / 2.0),   # per GPU
            "metrics_check": m,
        }
        if train is not None:
            line["train"] = train
        if int8 is not None:
            line["int8"] = int8
        if world == 1 and not args.no_cpu_baseline:
            v, n, per, kind = cpu_rate(os.cpu_count() or 1, budget_s=12.0)
            line["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": torch.get_num_threads(), "kind": kind,
                                    "sample": "%d sequential eval forwards of the B=256 batch (%.3f s each), scaled to S=100; %s" % (
                                        n, per, "unmodified reference modules (oracle/_ref)" if kind == "reference" else "oracle port")}
            if not args.no_gpu_eager:
                # context only (SURVEY 2.2 names cuDNN/cuBLAS on the GPU as the reference's own fast path): the SAME reference
                # modules on this B200 through PyTorch eager; not the product, not the baseline arm
                try:
                    vg, ng, perg, kindg = cpu_rate(os.cpu_count() or 1, budget_s=6.0, max_samples=100, device=str(dev))
                    if kindg == "reference":
                        line["reference_gpu_eager"] = {"value": vg, "unit": "images/s",
                                                       "sample": "%d eval forwards (%.4f s each) of the unmodified reference modules on this GPU (PyTorch eager, "
                                                                 "cuDNN/cuBLAS, torch's default conv precision), scaled to S=100; context only" % (ng, perg)}
                except Exception as e:
                    line["reference_gpu_eager"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
