#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_lrt_p4.py -q 2>&1 | tail -120 > gpurun_out/c4_lrt_tests_all.log
timeout 120 python scripts/dbg/wgrad_diag.py > gpurun_out/c4_wgrad_diag.txt 2>&1
timeout 120 python scripts/profile_train.py tf32 > gpurun_out/c4_profile_train.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-int8 > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/c4_suite.log
tail -8 gpurun_out/c4_lrt_tests_all.log; cat gpurun_out/c4_wgrad_diag.txt; tail -5 gpurun_out/c4_suite.log
