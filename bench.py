#!/usr/bin/env python
"""bench.py — headline benchmark of the stochastic-layer hot path.

Metric (BASELINE.json): ResNet-18 (24/48/96/192) Bayes-by-backprop MC-sampled images/sec at S=100
on synthetic CIFAR-shape data (B=256 x 3x32x32, random-init trained-like weights).

  python bench.py --gpus N --steps K --warmup W          our arm (one process per GPU under torchrun)
  python bench.py --impl reference ...                   the reference's CPU path (oracle port), rank 0 only

One "step" = one pass of the hot path over one batch: 256 images x 100 MC samples -> p-bar -> metrics.
N > 1: the 100 samples are sharded over ranks (global Philox sample index), one NCCL allreduce of
the [256,10] probability sums per step; total work per step is fixed => "scaling": "strong".
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
# mean dram__bytes_read.sum + dram__bytes_write.sum per conv launch of one 10-sample chunk (ncu --set full)
P4_DRAM_BYTES_PER_LAUNCH = 314.6e6
P4_TRAFFIC_SOURCE = "dram__bytes_read.sum + dram__bytes_write.sum, mean over the 17 umma_conv_p4_kernel launches of one 10-sample chunk, ncu --set full (profiles/r01_p4_kernels_ncu_full.csv: 314.6 MB), scaled linearly to this run's samples per chunk; the residual reads add to SURVEY's algorithmic figure, the shared borders subtract"
ACT_BYTES_PER_SAMPLE_IMAGE = 1.929e6      # fp32 NHWC activations in+out of the 21 stochastic layers


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


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


def cpu_port_rate(n_threads, budget_s=12.0, max_samples=40):
    """The reference's CPU path (oracle port: same torch-CPU operators, experiments/utils.py:342-355 loop)
    on a bounded sample: S' sequential eval forwards of the B=256 batch, scaled to S=100."""
    import torch
    import oracle.qbn_oracle as O
    torch.set_num_threads(n_threads)
    P = O.ResNetBBBParams(seed=1)
    x = torch.randn(B, 3, 32, 32, generator=torch.Generator().manual_seed(2))
    eps_fn = lambda name, shape: torch.empty(shape).normal_()  # noqa: E731
    with torch.no_grad():
        O.resnet_bbb_eval_forward(P, x, eps_fn)  # warm-up
        t0 = time.perf_counter()
        n = 0
        while n < max_samples and (time.perf_counter() - t0 < budget_s or n < 3):
            O.resnet_bbb_eval_forward(P, x, eps_fn)
            n += 1
        dt = time.perf_counter() - t0
    per_sample = dt / n
    return B / (per_sample * S), n, per_sample


def run_reference(args):
    """--impl reference: rank 0 times the CPU path with all host threads; other ranks exit 0."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    vals = []
    sp = None
    for _ in range(args.warmup if args.warmup < 1 else 1):
        cpu_port_rate(cores, budget_s=2.0, max_samples=3)
    for _ in range(max(1, args.steps)):
        v, n, per = cpu_port_rate(cores, budget_s=max(2.0, 60.0 / max(1, args.steps)), max_samples=5)
        vals.append(v)
        sp = (n, per)
    value = statistics.mean(vals)
    line = {
        "impl": "reference", "metric": "resnet18_bbb_mc_images_per_sec_S100", "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * B / value, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ResNet-18(24/48/96/192) BBB eval, B=256 x 3x32x32, S=100 MC samples", "global_batch": B, "samples": S},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": "each step = %d sequential eval forwards of the B=256 batch (%.3f s each) scaled x%d/%d to S=100; "
                                   "oracle port = the torch-CPU operators the reference calls" % (sp[0], sp[1], S, sp[0])},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


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
    from qbn_b200 import mc, metrics, noise, synthetic, zoo   # the GPU arm never imports oracle/ (only the cpu_baseline leg does)

    dev = torch.device("cuda", local_rank)
    P = synthetic.ResNetBBBParams(seed=1)
    model = zoo.resnet_from_params(P).to(dev).eval()
    noise.manual_seed(20261017)
    engine = mc.MCEngine(model, math_mode=args.math, chunk=args.chunk, chunk_max=args.chunk_max)
    x_host = torch.randn(B, 3, 32, 32, generator=torch.Generator().manual_seed(2)).pin_memory()
    t_host = torch.randint(0, K_CLASSES, (B,), generator=torch.Generator().manual_seed(3)).pin_memory()
    x_dev, t_dev = x_host.to(dev), t_host.to(dev)
    start, count = qdist.shard_range(S, rank, world)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2
    metric = metrics.ClassificationMetric(K_CLASSES, device=dev)

    def step_resident():
        psum = engine.predict_sum(x_dev, count, sample0=start)
        qdist.allreduce_prob_sums(psum)
        metric.update(psum, t_dev, scale=1.0 / S)
        return psum

    def step_e2e():
        xd = x_host.to(dev, non_blocking=True)
        td = t_host.to(dev, non_blocking=True)
        psum = engine.predict_sum(xd, count, sample0=start)
        qdist.allreduce_prob_sums(psum)
        metric.update(psum, td, scale=1.0 / S)
        probs = (psum / S).to("cpu", non_blocking=True)
        st = metric.state.to("cpu", non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return probs, st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- timed: K steps, each bracketed by CUDA events on the launching stream; L2 flushed between steps
    times = []
    engine.launches = 0
    for _ in range(args.steps):
        flush.fill_(1.0)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step_resident()
        e1.record()
        barrier()
        times.append(e0.elapsed_time(e1))
    launches = engine.launches + 2 * args.steps  # + metric kernel (+ allreduce)
    total_ms = torch.tensor([sum(times)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    # ---- e2e: host buffers, H2D + D2H inside the timed region (wall clock around synchronous steps, max over ranks)
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline of the dominant kernel (umma_conv_kernel<EVAL>): per-launch CUDA events on the launch stream
    roof = None
    if rank == 0:
        peaks = _peaks()
        evs = []
        orig, orig_s1 = mc.ops.conv_forward, mc.ops.conv_s1_forward

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

        def timed_s1(x, w, n, N, R, S_, *a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = orig_s1(x, w, n, N, R, S_, *a, **k)
            e1.record()
            H, W = x.shape[2] - (R - 1), x.shape[3] - (S_ - 1)       # algorithmic flops: interior pixels only
            evs.append((e0, e1, 2.0 * x.shape[0] * H * W * N * R * S_ * x.shape[1], 1))
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
        mc.ops.conv_forward, mc.ops.conv_s1_forward, mc.ops.conv_p4_forward = timed_conv, timed_s1, timed_p4
        mc.ops.conv_p4_shortcut_forward = timed_p4sc
        flush.fill_(1.0)
        torch.cuda.synchronize()
        engine.use_graph = False                      # per-launch events need the eager launch sequence (the timed loop replays a CUDA graph)
        engine.predict_sum(x_dev, count, sample0=start)
        engine.use_graph = True
        torch.cuda.synchronize()
        mc.ops.conv_forward, mc.ops.conv_s1_forward, mc.ops.conv_p4_forward = orig, orig_s1, orig_p4
        mc.ops.conv_p4_shortcut_forward = orig_p4sc
        um = [(a.elapsed_time(b), f) for a, b, f, m in evs if m >= 1]
        n_p4 = sum(1 for e in evs if e[3] == 2)
        if um:
            t_ms = sum(t for t, _ in um)
            fl = sum(f for _, f in um)
            ach = fl / (t_ms * 1e-3) / 1e12
            peak = peaks["bf16_tflops_sustained"] / 2.0
            alg_bytes = ACT_BYTES_PER_SAMPLE_IMAGE * B * count
            # with fp32 activations the eval path's arithmetic intensity (81 flop/B) is below the TF32 ridge
            # (~105 flop/B at the measured peaks), so HBM is the binding roofline; the tensor-pipe view is reported beside it.
            hbm = alg_bytes / (t_ms * 1e-3) / 1e9
            roof = {"kernel": "umma_conv_p4_kernel (%d launches) + umma_conv_kernel<EVAL> (%d) — tcgen05 kind::tf32" % (n_p4, len(um) - n_p4),
                    "bound": "hbm", "achieved": hbm, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm / peaks["hbm_gbs"],
                    "traffic": P4_DRAM_BYTES_PER_LAUNCH * (count / float(-(-count // (args.chunk_max or args.chunk)))) / 10.0,
                    "launches": len(um), "avg_launch_ms": t_ms / len(um),
                    "peak_source": "hbm_gbs of %s" % peaks["source"],
                    "traffic_source": P4_TRAFFIC_SOURCE,
                    "algorithmic_bytes_per_launch": alg_bytes / len(um),
                    "algorithmic_bytes_note": "1.929 MB per sample-image (SURVEY 8d: activations in + out of the 21 stochastic layers, fp32) x the "
                                              "sample-images of the step / conv launches",
                    "tensor_view": {"achieved_tflops": ach, "peak_tflops": peak, "frac": ach / peak,
                                    "peak_source": "1/2 x sustained bf16 of %s (TF32 peak not in MEASURED_PEAKS.json)" % peaks["source"]},
                    "share_of_step": t_ms / (total_ms / args.steps)}
    if rank == 0:
        value = B * args.steps / (total_ms * 1e-3)
        e2e_v = B * args.steps / e2e_s
        m = metric.compute()
        line = {
            "metric": "resnet18_bbb_mc_images_per_sec_S100", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "tf32" if args.math == "tf32" else "f32", "data": "synthetic",
            "config": {"workload": "ResNet-18(24/48/96/192) BBB eval, B=256 x 3x32x32, S=100 MC samples", "global_batch": B, "samples": S,
                       "parallelism": "mc-sample sharding x%d" % world, "chunk": args.chunk,
                       "l2": "256 MB flush between timed steps; per-layer activations (S x 25 MB) exceed the 126 MB L2"},
            "sample_images_per_sec": value * S,
            "e2e": {"value": e2e_v, "unit": "images/s", "h2d_bytes_per_step": x_host.numel() * 4 + t_host.numel() * 8,
                    "d2h_bytes_per_step": B * K_CLASSES * 4 + metric.state.numel() * 4},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof,
            "tensor_bound_frac_whole_step": (FLOP_PER_SAMPLE_IMAGE * B * S * args.steps / (total_ms * 1e-3) / 1e12) / (_peaks()["bf16_tflops_sustained"] / 2.0),
            "metrics_check": m,
        }
        if not args.no_cpu_baseline and world == 1:
            v, n, per = cpu_port_rate(os.cpu_count() or 1, budget_s=12.0)
            line["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": "%d sequential eval forwards of the B=256 batch (%.3f s each), scaled to S=100" % (n, per)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
