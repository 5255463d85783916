#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/profile_train_kernels.py > gpurun_out/c6_train_kernels.txt 2>&1
grep -A2 "launch order" gpurun_out/c6_train_kernels.txt
