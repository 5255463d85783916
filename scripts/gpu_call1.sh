#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lrt_p4.py -q -x 2>&1 | tail -40 > gpurun_out/c1_lrt_tests.log
timeout 300 python -m pytest tests/test_gpu_lrt_p4.py -q 2>&1 | tail -60 > gpurun_out/c1_lrt_tests_all.log
timeout 120 python scripts/profile_train.py tf32 > gpurun_out/c1_profile_train.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:umma_conv -c 44 -o /tmp/r02_p4 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-train --no-int8 --no-gpu-eager > gpurun_out/c1_ncu.log 2>&1
ncu -i /tmp/r02_p4.ncu-rep --page raw --csv > gpurun_out/r02_p4_raw.csv 2>/dev/null
ls -la /tmp/r02_p4.ncu-rep >> gpurun_out/c1_ncu.log
tail -5 gpurun_out/c1_lrt_tests.log
