"""Throughput of the other halves of BASELINE config 5 at B=256 (diagnostic: the models come from the reference's own lifecycle, so
this lives with the tests): the SGHMC-style 10-member int8 ensemble (Int8EnsembleEngine, every member on the planar kind::i8 kernel)
and the int8 MC-Dropout ResNet-18 at S=100 (Int8MCEngine, all samples of a chunk per launch).
Usage (GPU box): python tests/diag/diag_bench_c5_variants.py"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
import _ref_models as R  # noqa: E402


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    import __graft_entry__ as ge
    ge.build()
    from qbn_b200 import noise, quant_utils as qu
    from qbn_b200.mc_int8 import Int8EnsembleEngine, Int8MCEngine, make_int8_engine
    torch.set_num_threads(8)
    x = torch.randn(256, 3, 32, 32, generator=torch.Generator().manual_seed(2)).cuda()
    net, _ = R.sgld_ensemble(10)
    eng = Int8EnsembleEngine(qu.to_device_int8(R.clone(net), "cuda"))
    ms = timeit(lambda: eng.predict(x))
    print(json.dumps({"workload": "SGHMC 10-member int8 ResNet-18 ensemble eval, B=256 (config 5)", "engine": "Int8EnsembleEngine (planar kind::i8)",
                      "ms_per_batch": ms, "images_per_s": 256 / (ms * 1e-3), "member_images_per_s": 2560 / (ms * 1e-3)}))
    net, _ = R.mc_dropout_resnet()
    mine = qu.to_device_int8(R.clone(net), "cuda")
    noise.manual_seed(3)
    for eng in (make_int8_engine(mine, chunk=50), Int8MCEngine(mine, chunk=25)):
        ms = timeit(lambda: eng.predict(x, 100), 3)
        print(json.dumps({"workload": "int8 MC-Dropout ResNet-18 (p=0.15) eval, S=100, B=256 (config 5)", "engine": type(eng).__name__,
                          "ms_per_batch": ms, "images_per_s": 256 / (ms * 1e-3)}))


if __name__ == "__main__":
    main()
