#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/profile_train_kernels.py > gpurun_out/c5_train_kernels.txt 2>&1
timeout 200 python -m pytest tests/test_gpu_lrt_p4.py -q -k graphed 2>&1 | tail -5 > gpurun_out/c5_graph_test.log
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/c5_suite.log
cat gpurun_out/c5_train_kernels.txt | tail -50; tail -5 gpurun_out/c5_graph_test.log; tail -8 gpurun_out/c5_suite.log
