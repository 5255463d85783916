import torch, torch.nn as nn, torch.nn.functional as F
torch.manual_seed(0)
def rel(a, b): return "%.2e" % float((a - b).abs().max() / b.abs().max())
def run(fn, x, cl):
    xi = (x.contiguous(memory_format=torch.channels_last) if cl else x.clone()).requires_grad_(True)
    y = fn(xi)
    g = torch.randn(y.shape, generator=torch.Generator().manual_seed(1)).cuda()
    y.backward(g)
    return y.detach(), xi.grad.detach()
x = torch.randn(4, 192, 4, 4).cuda()
bn = nn.BatchNorm2d(192).cuda().train()
ops_ = {
    "bn_train": lambda t: bn(t),
    "relu": lambda t: F.relu(t),
    "avgpool4": lambda t: F.avg_pool2d(t, 4),
    "avgpool4+flatten+sum": lambda t: F.avg_pool2d(t, 4).reshape(4, -1) * 2.0,
    "add_relu": lambda t: F.relu(t + 0.5 * t.detach().flip(0)),
    "bn_add_relu_pool": lambda t: F.avg_pool2d(F.relu(bn(t) + t), 4).reshape(4, -1),
}
for name, fn in ops_.items():
    ya, ga = run(fn, x, False); yb, gb = run(fn, x, True)
    print("%-24s fwd %s  grad %s" % (name, rel(yb, ya), rel(gb, ga)))
# the actual pattern: relu output feeding avgpool, both CL
def block(t):
    r = F.relu(bn(t) + t)
    return F.avg_pool2d(r, 4).reshape(4, -1)
ya, ga = run(block, x, False); yb, gb = run(block, x, True)
print("block", rel(yb, ya), rel(gb, ga))
