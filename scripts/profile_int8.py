"""Per-launch timing of one chunk of the int8 engine: CUDA events around every libqbn entry point (and the torch glue
between them), so the split between the kind::i8 contractions, the weight sampler and the elementwise glue is visible.
Usage: python scripts/profile_int8.py [chunk=25] [tensor_cores=1]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def main():
    chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 25
    tensor_cores = bool(int(sys.argv[2])) if len(sys.argv) > 2 else True
    from bench_int8 import build_model
    net, x, _ = build_model(B=256)
    from qbn_b200 import _lib
    from qbn_b200.mc_int8 import Int8MCEngine, Int8PlanarEngine
    planar = len(sys.argv) > 3 and sys.argv[3] == "planar"
    eng = Int8PlanarEngine(net, chunk=chunk, use_graph=False) if planar else Int8MCEngine(net, chunk=chunk, tensor_cores=tensor_cores)
    eng.predict(x, chunk)                                   # warm-up (kernel attributes, allocator)
    rows, real_call = [], _lib.call

    def timed_call(name, *args):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = real_call(name, *args)
        e1.record()
        rows.append((name, e0, e1))
        return out

    _lib.call = timed_call                                  # ops.py resolves `_lib.call` at call time
    try:
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        eng.predict(x, chunk)
        t1.record()
        torch.cuda.synchronize()
    finally:
        _lib.call = real_call
    total = t0.elapsed_time(t1)
    per, count = {}, {}
    for name, e0, e1 in rows:
        per[name] = per.get(name, 0.0) + e0.elapsed_time(e1)
        count[name] = count.get(name, 0) + 1
    print("int8 ResNet-18, B=256, one chunk of %d samples, tensor_cores=%s: %.3f ms total (GPU events)" % (chunk, tensor_cores, total))
    inside = sum(per.values())
    for name in sorted(per, key=per.get, reverse=True):
        print("  %-26s %4d launches %9.3f ms  %5.1f %%" % (name, count[name], per[name], 100 * per[name] / total))
    print("  %-26s %4s          %9.3f ms  %5.1f %%   (permutes, clamps, dequantise, softmax, allocator)" % ("torch glue between calls", "", total - inside, 100 * (total - inside) / total))
    convs = [(e0.elapsed_time(e1)) for name, e0, e1 in rows if name in ("qbn_i8_conv_fwd", "qbn_i8_conv_p16_fwd")]
    print("  per-layer conv ms:", " ".join("%.3f" % v for v in convs))
    if planar:
        print("  per-layer steps  :", " ".join("%s:%dx%d/%d" % (st.name.replace("layers.", "L"), st.C, st.N, st.stride) for st in eng.steps))


if __name__ == "__main__":
    main()
