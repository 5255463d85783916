#!/bin/bash
# One GPU call that re-establishes the state of the tree on a fresh B200: parity suite, the full bench line (headline + training +
# int8 legs + CPU baseline), the reference arm, the per-kernel profile of the graphed training step.  Everything lands in
# gpurun_out/ (scratch).     gpurun --timeout 900 -- bash scripts/round_start.sh
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/rs_suite.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/rs_bench.json 2> gpurun_out/rs_bench.err; tail -c 1500 gpurun_out/rs_bench.json
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/rs_bench_ref.json 2> gpurun_out/rs_bench_ref.err; tail -c 400 gpurun_out/rs_bench_ref.json
timeout 120 python scripts/profile_train_kernels.py > gpurun_out/rs_train_kernels.txt 2>&1; head -12 gpurun_out/rs_train_kernels.txt | cut -c1-160
timeout 90 python scripts/bench_variants.py > gpurun_out/rs_variants.json 2>&1; tail -5 gpurun_out/rs_variants.json
