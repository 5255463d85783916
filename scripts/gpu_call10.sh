#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "i8 or int8 or quant" 2>&1 | tail -8 > gpurun_out/c10_i8_tests.log
timeout 200 python scripts/bench_int8.py 2>&1 | tail -1 > gpurun_out/c10_int8_bench.json
tail -4 gpurun_out/c10_i8_tests.log; cat gpurun_out/c10_int8_bench.json
