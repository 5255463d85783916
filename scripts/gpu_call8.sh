#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_lrt_p4.py tests/test_gpu_models.py -q 2>&1 | tail -8 > gpurun_out/c8_tests.log
QBN_PDL=0 timeout 200 python scripts/profile_train_kernels.py > gpurun_out/c8_train_kernels_pdl0.txt 2>&1
QBN_PDL=1 timeout 200 python scripts/profile_train_kernels.py > gpurun_out/c8_train_kernels_pdl1.txt 2>&1
tail -4 gpurun_out/c8_tests.log; grep "graphed step" gpurun_out/c8_train_kernels_pdl*.txt; sed -n 5,16p gpurun_out/c8_train_kernels_pdl0.txt | cut -c1-420
