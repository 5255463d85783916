"""Config 4 (BASELINE.json): CIFAR-shape ResNet-18 Bayes-by-backprop LRT training step, data-parallel.
One process per GPU (torchrun) or a single process; B=256 per GPU (weak scaling), Adam lr 1e-3, gamma=.01,
n_batches=176.  A step = LRT forward, KL, ELBO, backward, NaN scrub, gradient allreduce, Adam.
Prints one JSON line (secondary metric; the headline bench is bench.py)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--math", default="tf32")
ap.add_argument("--batch", type=int, default=256)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
import __graft_entry__ as ge
if rank == 0:
    ge.build()
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dist.barrier()
from qbn_b200 import synthetic as O      # seeded parameter containers
from qbn_b200 import config, noise, zoo
from qbn_b200 import dist as qdist
config.set_math_mode(args.math)
dev = torch.device("cuda", local)
model = zoo.resnet_from_params(O.ResNetBBBParams(seed=1)).to(dev).train()
if world > 1:
    qdist.broadcast_parameters(model)
noise.manual_seed(1234 + rank)      # independent epsilon substreams per replica
params = [p for p in model.parameters() if p.requires_grad]
opt = torch.optim.Adam(params, lr=1e-3)
B = args.batch
g = torch.Generator().manual_seed(5 + rank)
x = torch.randn(B, 3, 32, 32, generator=g).to(dev)
t = torch.randint(0, 10, (B,), generator=g).to(dev)

def step():
    opt.zero_grad(set_to_none=True)
    y = model(x)
    kl = model.get_kl_divergence()
    loss = torch.nn.functional.nll_loss(torch.log(y + 1e-8), t) + 0.01 * kl / (B * world * 176)
    loss.backward()
    if world > 1:
        qdist.allreduce_gradients(params)
    else:
        qdist.scrub_nan_grads(params)
    opt.step()
    return loss

for _ in range(max(3, args.warmup)):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    ms_step = float(ms) / args.steps
    flop = 9.396e8 * B * world        # F_train per image
    print(json.dumps({"metric": "resnet18_bbb_lrt_train_images_per_sec", "value": B * world / (ms_step * 1e-3), "unit": "images/s",
                      "n_gpus": world, "steps": args.steps, "ms_per_step": ms_step, "scaling": "weak", "dtype": args.math,
                      "config": {"workload": "ResNet-18 BBB LRT training step, B=256 per GPU, Adam, KL + NLL", "parallelism": "dp%d" % world},
                      "tflops": flop / (ms_step * 1e-3) / 1e12, "loss": float(loss)}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
