#!/bin/bash
# the whole GPU suite under compute-sanitizer memcheck (bounded), plus racecheck on one planar-conv and one LRT case
mkdir -p gpurun_out
timeout 560 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/san_all.log python -m pytest tests -m gpu -q --deselect tests/test_bench_contract.py > gpurun_out/san_all.out 2>&1
echo "memcheck suite: rc=$? $(tail -1 gpurun_out/san_all.out) | $(tail -1 gpurun_out/san_all.log)"
