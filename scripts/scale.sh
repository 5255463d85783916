#!/bin/bash
# 1/2/4/8-GPU scaling of the bench (headline + training + int8 legs) on one box: gpurun --gpus 8 -- bash scripts/scale.sh
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager 2>gpurun_out/scale_n$n.err | tail -1 > gpurun_out/scale_n$n.json
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager 2>gpurun_out/scale_n$n.err | tail -1 > gpurun_out/scale_n$n.json
  fi
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/scale_n$n.json").read())
    print("N=$n value %.0f e2e %.0f frac %.3f | train %.0f img/s %.2f ms (%s) | int8 %.0f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["train"]["value"], d["train"]["ms_per_step"], d["train"]["mode"][:12], d["int8"]["value"]))
except Exception as e:
    print("N=$n failed", e)
PY
done
