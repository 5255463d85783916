import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch, torch.nn as nn
import torch.nn.functional as F
import oracle.qbn_oracle as O
import __graft_entry__ as ge
ge.build()
from qbn_b200 import zoo, noise, config, ops
from qbn_b200.stochastic.bbb.conv import Conv2d
from qbn_b200.stochastic.bbb.linear import Linear
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
config.set_math_mode("fp32")
def rel(a, b):
    a = a.detach().float().cpu(); b = (b.detach().cpu() if torch.is_tensor(b) else torch.as_tensor(np.asarray(b))).float()
    return "%.2e" % float((a-b).abs().max()/b.abs().max())
class Twin(nn.Module):
    def __init__(self, m):
        super().__init__(); self.m = m
        self.weight, self.std, self.bias = m.weight, m.std, m.bias
    def forward(self, x):
        eps = noise.pop_injected()
        if isinstance(self.m, Conv2d):
            mean = F.conv2d(x, self.weight, None, self.m.stride, self.m.padding)
            var = F.conv2d(x * x, F.softplus(self.std) ** 2, None, self.m.stride, self.m.padding)
        else:
            mean = x @ self.weight.t(); var = (x * x) @ (F.softplus(self.std) ** 2).t()
        return mean + torch.sqrt(1e-8 + var) * eps
    def get_kl_divergence(self): return self.m.get_kl_divergence()
def swap(mod):
    for n, c in list(mod.named_children()):
        if isinstance(c, (Conv2d, Linear)): setattr(mod, n, Twin(c))
        elif isinstance(c, nn.ModuleList):
            for i, cc in enumerate(c):
                if isinstance(cc, (Conv2d, Linear)): c[i] = Twin(cc)
                else: swap(cc)
        else: swap(c)
g = np.load("tests/golden/resnet.npz")
P = O.ResNetBBBParams(seed=21)
x = torch.randn(4, 3, 32, 32, generator=torch.Generator().manual_seed(22))
tgt = torch.randint(0, 10, (4,), generator=torch.Generator().manual_seed(23)).cuda()
def run(twin, use_kl=True, cl=True):
    net = zoo.resnet_from_params(P).cuda()
    m2 = copy.deepcopy(net).eval(); order = []
    names = {mod: name for name, mod in m2.named_modules()}
    for mod in m2.modules():
        if isinstance(mod, (Conv2d, Linear)):
            mod.register_forward_hook(lambda mod, i, o: order.append((names[mod], tuple(o.shape))))
    with torch.no_grad(): m2(x.cuda())
    eps = O.replay_noise(710, [o[1] for o in order])
    kl_mods = [m for m in net.modules() if isinstance(m, (Conv2d, Linear))]
    if twin: swap(net)
    net.train()
    grads_out = {}
    with noise.inject([e.cuda() for e in eps]):
        y = net(x.cuda())
    kl = sum(m.get_kl_divergence() for m in kl_mods)
    loss = F.nll_loss(torch.log(y + 1e-8), tgt) + (0.01 * kl / (4 * 176) if use_kl else 0.0)
    loss.backward()
    return y, {n: p.grad for n, p in net.named_parameters() if p.grad is not None}
y1, g1 = run(False); y2, g2 = run(True)
print("fwd ours vs twin", rel(y1, y2), "ours vs golden", rel(y1, g["y_train"]))
keys = ["layers.9", "layers.6.1.stem.3", "layers.6.0.shortcut.0", "layers.5.0.shortcut.0", "layers.4.0.stem.0", "layers.3.0.stem.0", "layers.0"]
for k in keys:
    for suf in (".weight", ".std"):
        a = g1[k + suf]; b = g2.get(k + suf, g2.get(k + ".m" + suf))
        line = "%-28s ours-vs-twin %s" % (k + suf, rel(a, b))
        if "g." + k + suf in g.files: line += "  ours-vs-golden %s  twin-vs-golden %s" % (rel(a, g["g." + k + suf]), rel(b, g["g." + k + suf]))
        print(line)
print("bn layers.1.weight ours-vs-twin", rel(g1["layers.1.weight"], g2["layers.1.weight"]), "ours-vs-golden", rel(g1["layers.1.weight"], g["g.layers.1.weight"]), "twin-vs-golden", rel(g2["layers.1.weight"], g["g.layers.1.weight"]))

# ---- finer: capture per-layer (x, eps, grad_out) in both runs ------------------------------------
print("==== per-layer grad_output / weight-grad comparison (backward order) ====")
def run2(twin):
    net = zoo.resnet_from_params(P).cuda()
    m2 = copy.deepcopy(net).eval(); order = []
    names = {mod: name for name, mod in m2.named_modules()}
    for mod in m2.modules():
        if isinstance(mod, (Conv2d, Linear)):
            mod.register_forward_hook(lambda mod, i, o: order.append((names[mod], tuple(o.shape))))
    with torch.no_grad(): m2(x.cuda())
    eps = O.replay_noise(710, [o[1] for o in order])
    if twin: swap(net)
    net.train()
    cap = {}
    lname = {mod: name for name, mod in net.named_modules()}
    def fh(mod, inp, out):
        n = lname[mod]
        cap[n] = {"x": inp[0].detach().clone(), "out": out.detach().clone()}
        out.register_hook(lambda g, n=n: cap[n].__setitem__("g", g.detach().clone()))
    for mod in net.modules():
        if isinstance(mod, (Conv2d, Linear, Twin)) and not (twin and isinstance(mod, (Conv2d, Linear))):
            mod.register_forward_hook(fh)
    with noise.inject([e.cuda() for e in eps]):
        y = net(x.cuda())
    loss = F.nll_loss(torch.log(y + 1e-8), tgt)
    loss.backward()
    grads = {n: p.grad for n, p in net.named_parameters() if p.grad is not None}
    return cap, grads, dict(zip([o[0] for o in order], eps))
c1, g1, eps_by = run2(False); c2, g2, _ = run2(True)
for n in reversed(list(c1.keys())):
    a, b = c1[n], c2[n]
    print("%-24s x %s out %s g_out %s | dW %s dRho %s | g strides %s x strides %s" % (n, rel(a["x"], b["x"]), rel(a["out"], b["out"]), rel(a["g"], b["g"]),
          rel(g1[n + ".weight"], g2[n + ".weight"]), rel(g1[n + ".std"], g2[n + ".std"]), tuple(a["g"].stride()), tuple(a["x"].stride())))
# reproduce the last conv standalone with captured tensors
n = "layers.6.1.stem.3"
mod = dict(zoo.resnet_from_params(P).named_modules())[n]
xx, gg, ee = c1[n]["x"], c1[n]["g"], eps_by[n].cuda()
mu, rho = P.convs[n]
xr = xx.clone().requires_grad_(True); mur = mu.cuda().requires_grad_(True); rhor = rho.cuda().requires_grad_(True)
yy = ops.LRTFunction.apply(xr, mur, rhor, None, 1, 1, 1, ee, (0, 0, 0), 0, False, None)
yy.backward(gg)
yo, so = O.lrt_conv_fwd(xx.cpu(), mu, rho, None, ee.cpu(), 1, 1)
dx, dmu, drho, _ = O.lrt_conv_bwd(xx.cpu(), mu, rho, ee.cpu(), so, gg.cpu(), 1, 1)
print("standalone", n, "fwd", rel(yy, yo), "dx", rel(xr.grad, dx), "dmu", rel(mur.grad, dmu), "drho", rel(rhor.grad, drho), "in-net dmu vs oracle", rel(g1[n + ".weight"], dmu))
print("min std", float(so.min()), "frac x==0", float((xx == 0).float().mean()))
