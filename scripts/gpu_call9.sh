#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_lrt_p4.py tests/test_gpu_models.py -q 2>&1 | tail -8 > gpurun_out/c9_tests.log
timeout 200 python scripts/profile_train_kernels.py > gpurun_out/c9_train_kernels.txt 2>&1
tail -4 gpurun_out/c9_tests.log; grep "graphed step" gpurun_out/c9_train_kernels.txt; sed -n 5,22p gpurun_out/c9_train_kernels.txt | cut -c1-250
