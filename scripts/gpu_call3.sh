#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/dbg/mn_probe.py > gpurun_out/c3_mn_probe.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_lrt_p4.py -q -k "graphed" 2>&1 | tail -30 > gpurun_out/c3_graph_test.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-int8 > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err
tail -50 gpurun_out/c3_mn_probe.txt
