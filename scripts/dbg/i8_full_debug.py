import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pathlib
from test_gpu_i8_full_resnet import full_int8_resnet, replay_eps
from qbn_b200 import noise
from qbn_b200.mc_int8 import Int8PlanarEngine
gd = pathlib.Path(ROOT) / "tests" / "golden"
g = np.load(gd / "resnet_int8_full.npz")
net = full_int8_resnet(gd)
x = torch.as_tensor(g["x"]).cuda()
mods = dict(net.named_modules())
names = [str(n) for n in g["q_names"]]
inj = [[e.cuda() for e in replay_eps(net, g, fi)] for fi in (0, 1)]
ref = []
for fi in (0, 1):
    seen = {}
    hooks = [mods[n].register_forward_hook(lambda m, i, o, n=n: seen.__setitem__(n, o.q.clone())) for n in names]
    with torch.no_grad(), noise.inject(list(inj[fi])):
        net(x)
    for h in hooks: h.remove()
    ref.append(seen)
eng = Int8PlanarEngine(net, chunk=2, use_graph=False)
eng.trace = {}
eng.predict_sum(x, 2, injected=inj)
B = x.shape[0]
for st in eng.steps:
    ints = eng.trace[st.name][0]
    if st.residual is not None:
        continue
    for fi in (0, 1):
        mine = ints[fi * B:(fi + 1) * B].cpu().numpy().astype(int)
        want = ref[fi][st.name].cpu().numpy().astype(int)
        bad = mine != want
        print(st.name, "sample", fi, "stride", st.stride, "k", st.ksize, "mismatch", int(bad.sum()), "of", bad.size)
        if bad.any():
            idx = np.argwhere(bad)
            print("  first bad (b,c,h,w):", idx[:8].tolist(), "mine", mine[bad][:8], "want", want[bad][:8])
            print("  bad per image", bad.sum(axis=(1, 2, 3)), "per channel (first 16)", bad.sum(axis=(0, 2, 3))[:16])
            print("  bad per row h", bad.sum(axis=(0, 1, 3)))
