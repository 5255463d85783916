import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import torch.nn.functional as F
import oracle.qbn_oracle as O
import __graft_entry__ as ge
ge.build()
from qbn_b200 import zoo, noise, config, ops
config.set_math_mode("fp32")
def rel(a, b):
    a = a.detach().float().cpu(); b = torch.as_tensor(np.asarray(b)).float()
    return float((a-b).abs().max()), float((a-b).abs().max()/b.abs().max())
G = lambda n: np.load(os.path.join("tests/golden", n + ".npz"))
# ---- LeNet eval layer by layer
P = O.LeNetBBBParams(seed=31)
x = torch.rand(4, 1, 28, 28, generator=torch.Generator().manual_seed(32))
nz = O.replay_noise(800, [p[1] for p in P.noise_plan()])
print("noise checksum", [float(t.double().sum()) for t in nz])
net = zoo.lenet_from_params(P).cuda().eval()
h_ref = x; h = x.cuda()
names = ["layers.0", "layers.2", "layers.5", "layers.7"]
i = 0
for li, layer in enumerate(net.layers):
    if hasattr(layer, "std"):
        mu, rho = P.layers[names[i]]
        with noise.inject([nz[i].cuda()]):
            h = layer(h)
        if mu.dim() == 4:
            h_ref = O.eval_conv_fwd(h_ref, mu, rho, None, nz[i], 1, 2, 1)
        else:
            h_ref = O.eval_linear_fwd(h_ref, mu, rho, None, nz[i])
        i += 1
    else:
        h = layer(h)
        h_ref = layer(h_ref)
    print("lenet layer", li, type(layer).__name__, "abs/rel err", rel(h, h_ref), "max|ref|", float(h_ref.abs().max()))
g = G("lenet")
print("final vs golden", rel(F.softmax(h, -1), g["y_eval0"]), "oracle vs golden", rel(F.softmax(h_ref, -1), g["y_eval0"]))
# ---- ResNet train grads per key
g = G("resnet")
P = O.ResNetBBBParams(seed=21)
x = torch.randn(4, 3, 32, 32, generator=torch.Generator().manual_seed(22))
net = zoo.resnet_from_params(P).cuda()
from qbn_b200.stochastic.bbb.conv import Conv2d
from qbn_b200.stochastic.bbb.linear import Linear
m2 = copy.deepcopy(net).eval(); order = []; hooks = []
names = {mod: name for name, mod in m2.named_modules()}
for mod in m2.modules():
    if isinstance(mod, (Conv2d, Linear)):
        hooks.append(mod.register_forward_hook(lambda mod, i, o: order.append((names[mod], tuple(o.shape)))))
with torch.no_grad(): m2(x.cuda())
net.train()
tgt = torch.randint(0, 10, (4,), generator=torch.Generator().manual_seed(23))
eps = O.replay_noise(710, [o[1] for o in order])
with noise.inject([e.cuda() for e in eps]):
    y = net(x.cuda())
print("train fwd", rel(y, g["y_train"]))
loss = F.nll_loss(torch.log(y + 1e-8), tgt.cuda()) + 0.01 * net.get_kl_divergence() / (4 * 176)
print("loss", float(loss), float(g["loss"]))
loss.backward()
sd = dict(net.named_parameters())
for key in ("layers.9.weight", "layers.9.std", "layers.5.0.shortcut.0.weight", "layers.5.0.shortcut.0.std", "layers.1.weight", "layers.0.weight", "layers.0.std"):
    print("grad", key, rel(sd[key].grad, g["g." + key]))
# ---- per-layer backward check at ResNet sizes vs oracle
torch.manual_seed(0)
for (B, C, H, N, k, s, p) in [(4, 3, 32, 24, 3, 1, 1), (4, 24, 32, 24, 3, 1, 1), (4, 24, 32, 48, 3, 2, 1), (4, 48, 16, 48, 3, 1, 1), (4, 96, 8, 192, 3, 2, 1), (4, 192, 4, 192, 3, 1, 1), (4, 24, 32, 48, 1, 2, 0)]:
    xx = torch.randn(B, C, H, H); mu = torch.randn(N, C, k, k) * 0.1; rho = torch.empty(N, C, k, k).uniform_(-5, -2)
    yo, so = O.lrt_conv_fwd(xx, mu, rho, None, torch.zeros(1), s, p)
    eps = torch.randn_like(yo); gout = torch.randn_like(yo)
    yo, so = O.lrt_conv_fwd(xx, mu, rho, None, eps, s, p)
    dx, dmu, drho, _ = O.lrt_conv_bwd(xx, mu, rho, eps, so, gout, s, p)
    xr = xx.cuda().requires_grad_(True); mur = mu.cuda().requires_grad_(True); rhor = rho.cuda().requires_grad_(True)
    yy = ops.LRTFunction.apply(xr, mur, rhor, None, s, p, 1, eps.cuda(), (0, 0, 0), 0, False, None)
    yy.backward(gout.cuda())
    print("layer", (B, C, H, N, k, s, p), "fwd", rel(yy, yo), "dx", rel(xr.grad, dx), "dmu", rel(mur.grad, dmu), "drho", rel(rhor.grad, drho))
