"""Int8 ResNet-18 BBB (A7/W8), S=100 MC samples, B=256 — SURVEY §8d config C5 on one GPU, through Int8MCEngine.

Synthetic trained-like state (mu ~ N(0, 1/fan_in), rho ~ U(-6,-3), random BatchNorm statistics), the full lifecycle on the
device (prepare_model -> one train + one eval forward to calibrate every observer -> convert), then the sample-batched
engine timed with CUDA events.  Writes one JSON object to gpurun_out/int8_bench.json (and prints it)."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build_model(B=256, seed=3):
    import __graft_entry__ as ge
    ge.build()
    from qbn_b200 import noise, quant_utils as qu, zoo
    args = zoo.Args(sigma_prior=0.05, model="conv_resnet_bbb", q=True, at=True, activation_precision=7, weight_precision=8)
    net = zoo.ConvNetwork_ResNet([1, 3, 32, 32], 10, True, args)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in net.modules():
            if hasattr(m, "std") and hasattr(m, "weight"):
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) / m.weight[0].numel() ** 0.5)
                m.std.uniform_(-6.0, -3.0, generator=g)
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.1, generator=g)
                m.running_var.uniform_(0.5, 1.5, generator=g)
                m.weight.uniform_(0.5, 1.5, generator=g)
                m.bias.normal_(0, 0.1, generator=g)
    net.train()
    qu.prepare_model(net, args)
    if not torch.cuda.is_available():
        return net, None, args
    net = net.cuda()
    x = torch.randn(B, 3, 32, 32, generator=g).cuda()
    noise.manual_seed(11)
    net(x)                                               # QAT train forward: first observer update
    net.eval()
    with torch.no_grad():
        net(x)                                           # eval forward: calibrates add_weight / mul_noise
    qu.convert(net)
    return net.eval(), x, args


def time_engine(engine, x, samples, iters):
    engine.predict(x, samples)                           # warm-up
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(iters):
        p = engine.predict(x, samples)
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / iters, p


def main():
    from qbn_b200.mc_int8 import Int8MCEngine
    B, S = 256, 100
    wall = time.time()
    net, x, args = build_model(B)
    if x is None:
        print("no GPU: built and prepared the model only")
        return
    out = {"workload": "resnet18_bbb_int8_a7w8_S100_B256", "B": B, "samples": S, "lifecycle_s": round(time.time() - wall, 2)}
    results = {}
    from qbn_b200.mc_int8 import Int8PlanarEngine
    import sys as _sys
    chunks = [int(c) for c in os.environ.get("QBN_I8_CHUNKS", "50").split(",")]
    variants = [("planar_chunk%d" % c, None, c, 5) for c in chunks] + [("tcgen05_i8", True, 25, 3)]
    if "--imad" in _sys.argv:
        variants.append(("imad", False, 25, 1))
    for name, tc, chunk, iters in variants:
        try:
            eng = Int8PlanarEngine(net, chunk=chunk) if tc is None else Int8MCEngine(net, chunk=chunk, tensor_cores=tc)
            ms, p = time_engine(eng, x, S, iters)
            results[name] = p
            out[name] = {"ms_per_batch": round(ms, 3), "images_per_s": round(B / ms * 1e3, 1), "chunk": chunk,
                         "finite": bool(torch.isfinite(p).all()), "row_sum_err": float((p.sum(-1) - 1).abs().max())}
        except Exception as e:                            # keep the other variant's number
            out[name] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
    keys = list(results)
    if len(keys) >= 2:
        out["max_abs_diff_between_paths"] = max(float((results[keys[0]] - results[k]).abs().max()) for k in keys[1:])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "int8_bench.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
