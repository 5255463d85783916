import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "diag2.py")).read().split("y1, g1 = run(False)")[0])
def run3(twin):
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
    cap = {}; seq = []
    lname = {mod: name for name, mod in net.named_modules()}
    def fh(mod, inp, out):
        n = lname[mod] + ":" + type(mod).__name__
        if not torch.is_tensor(out) or not out.requires_grad: return
        seq.append(n)
        out.register_hook(lambda g, n=n: cap.__setitem__(n, g.detach().clone()))
    for mod in net.modules():
        if len(list(mod.children())) == 0 or isinstance(mod, Twin) or type(mod).__name__ in ("BasicBlock", "Add"):
            if twin and isinstance(mod, (Conv2d, Linear)): continue
            mod.register_forward_hook(fh)
    with noise.inject([e.cuda() for e in eps]):
        y = net(x.cuda())
    loss = F.nll_loss(torch.log(y + 1e-8), tgt)
    loss.backward()
    return cap, seq
c1, s1 = run3(False); c2, s2 = run3(True)
s2n = [n.replace(":Twin", "") for n in s2]
for n in reversed(s1[-14:]):
    base = n.split(":")[0]
    m = [k for k in c2 if k.split(":")[0] == base and (k.split(":")[1] == n.split(":")[1] or k.endswith(":Twin"))]
    if n in c1 and m:
        print("%-40s grad_out ours-vs-twin %s  strides %s vs %s" % (n, rel(c1[n], c2[m[0]]), tuple(c1[n].stride()), tuple(c2[m[0]].stride())))
# direct fc-shaped LRT backward check vs oracle
torch.manual_seed(3)
for (Bb, K, N) in [(4, 192, 10), (4, 192, 16), (8, 192, 10), (64, 100, 100), (4, 64, 10), (4, 65, 10), (4, 128, 10)]:
    xx = torch.rand(Bb, K); mu = torch.randn(N, K) * 0.1; rho = torch.empty(N, K).uniform_(-6, -4); eps = torch.randn(Bb, N); gout = torch.randn(Bb, N)
    yo, so = O.lrt_linear_fwd(xx, mu, rho, None, eps)
    dx, dmu, drho, _ = O.lrt_linear_bwd(xx, mu, rho, eps, so, gout)
    xr = xx.cuda().requires_grad_(True); mur = mu.cuda().requires_grad_(True); rhor = rho.cuda().requires_grad_(True)
    yy = ops.LRTFunction.apply(xr, mur, rhor, None, 1, 0, 1, eps.cuda(), (0, 0, 0), 0, False, None)
    yy.backward(gout.cuda())
    print("linear", (Bb, K, N), "fwd", rel(yy, yo), "dx", rel(xr.grad, dx), "dmu", rel(mur.grad, dmu), "drho", rel(rhor.grad, drho))
