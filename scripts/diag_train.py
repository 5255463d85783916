import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
from qbn_b200 import synthetic as O
from qbn_b200 import config, noise, zoo, ops
config.set_math_mode(sys.argv[1] if len(sys.argv) > 1 else "tf32")
model = zoo.resnet_from_params(O.ResNetBBBParams(seed=1)).cuda().train()
params = [p for p in model.parameters() if p.requires_grad]
opt = torch.optim.Adam(params, lr=1e-3)
B = 256
x = torch.randn(B, 3, 32, 32).cuda(); t = torch.randint(0, 10, (B,)).cuda()
def ev(): return torch.cuda.Event(enable_timing=True)
for it in range(6):
    e = [ev() for _ in range(5)]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e[0].record(); opt.zero_grad(set_to_none=True); y = model(x); e[1].record()
    kl = model.get_kl_divergence(); loss = torch.nn.functional.nll_loss(torch.log(y + 1e-8), t) + 0.01 * kl / (B * 176); e[2].record()
    loss.backward(); e[3].record(); opt.step(); e[4].record()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    if it >= 3:
        print("fwd %.2f  kl+loss %.2f  bwd %.2f  adam %.2f ms | cpu issue %.2f ms, total wall %.2f ms" % (e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3]), e[3].elapsed_time(e[4]), (t1 - t0) * 1e3, (t2 - t0) * 1e3))
# per-op kernel time with the torch profiler
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    opt.zero_grad(set_to_none=True); y = model(x); kl = model.get_kl_divergence()
    loss = torch.nn.functional.nll_loss(torch.log(y + 1e-8), t) + 0.01 * kl / (B * 176); loss.backward(); opt.step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
