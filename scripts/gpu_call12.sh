#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/c12_suite.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c12_bench_ref.json 2> gpurun_out/c12_bench_ref.err
tail -3 gpurun_out/c12_suite.log; tail -c 600 gpurun_out/c12_bench_ref.json
