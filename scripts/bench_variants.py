"""Secondary workloads (BASELINE.json configs 2 and 5, float): S=100 Monte-Carlo eval throughput of the other model
families through the same engine.  One JSON line per workload (not the headline; bench.py is)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
from qbn_b200 import synthetic as O      # seeded parameter containers
from qbn_b200 import mc, noise, zoo

def run(name, net, x, S=100, chunk=10, steps=5, warmup=3):
    eng = mc.MCEngine(net, math_mode="tf32", chunk=chunk)
    noise.manual_seed(1)
    flush = torch.empty(64 * 1024 * 1024, device="cuda")
    for _ in range(warmup):
        eng.predict_sum(x, S)
    torch.cuda.synchronize()
    ms = 0.0
    for _ in range(steps):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); p = eng.predict_sum(x, S); e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    ms /= steps
    print(json.dumps({"workload": name, "metric": "mc_images_per_sec_S100", "value": x.shape[0] / (ms * 1e-3), "unit": "images/s",
                      "ms_per_step": ms, "batch": x.shape[0], "samples": S, "dtype": "tf32", "prob_row_sum": float(p.sum() / (S * x.shape[0]))}))

P = O.ResNetBBBParams(seed=1)
x = torch.randn(256, 3, 32, 32, generator=torch.Generator().manual_seed(2)).cuda()
run("ResNet-18 MC-Dropout p=0.15 (config 5, float)", zoo.resnet_mc_from_params(P, 0.15, state_dict=O.resnet_mc_state_dict(P)).cuda().eval(), x)
run("ResNet-18 BBB (headline, for comparison)", zoo.resnet_from_params(P).cuda().eval(), x)
PL = O.LeNetBBBParams(seed=3)
xl = torch.rand(256, 1, 28, 28, generator=torch.Generator().manual_seed(4)).cuda()
run("LeNet MC-Dropout p=0.2 (config 2)", zoo.lenet_mc_from_params(PL, 0.2).cuda().eval(), xl)
run("LeNet BBB", zoo.lenet_from_params(PL).cuda().eval(), xl)
