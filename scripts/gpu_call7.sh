#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_lrt_p4.py -q 2>&1 | tail -40 > gpurun_out/c7_lrt_tests.log
timeout 200 python scripts/profile_train_kernels.py > gpurun_out/c7_train_kernels.txt 2>&1
tail -6 gpurun_out/c7_lrt_tests.log; head -14 gpurun_out/c7_train_kernels.txt | cut -c1-150
