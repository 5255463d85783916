#!/bin/bash
# 1/2/4/8-GPU scaling of the headline bench on one box (run under `gpurun --gpus 8`)
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps 10 --warmup 3 2>&1 | tail -1
  fi
done
