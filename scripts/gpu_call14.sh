#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_lrt_p4.py tests/test_gpu_models.py tests/test_gpu_reference_callsites.py -q 2>&1 | tail -8 > gpurun_out/c14_tests.log
timeout 200 python scripts/profile_train_kernels.py > gpurun_out/c14_train_kernels.txt 2>&1
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-train --no-int8 --chunk 100 > gpurun_out/c14_bench_chunk100.json 2>/dev/null
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-train --no-int8 --chunk 34 > gpurun_out/c14_bench_chunk34.json 2>/dev/null
tail -4 gpurun_out/c14_tests.log; grep "graphed step" gpurun_out/c14_train_kernels.txt; sed -n 5,16p gpurun_out/c14_train_kernels.txt | cut -c1-150
python -c "
import json
for c in (100, 34):
    d = json.loads(open('gpurun_out/c14_bench_chunk%d.json' % c).read().strip().splitlines()[-1]); print('chunk', c, d['value'], d['ms_per_step'])"
