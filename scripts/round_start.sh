#!/bin/bash
# One GPU call that re-establishes the state of the tree on a fresh B200: parity suite, headline bench, the other families,
# the training step and the per-entry-point split of the int8 engine.  Everything lands in gpurun_out/ (scratch).
#   gpurun --timeout 600 -- bash scripts/round_start.sh
mkdir -p gpurun_out
timeout 120 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/rs_suite.log
timeout 150 python bench.py --steps 10 --warmup 3 > gpurun_out/rs_bench.json 2> gpurun_out/rs_bench.err; tail -c 1200 gpurun_out/rs_bench.json
timeout 90 python scripts/bench_variants.py > gpurun_out/rs_variants.json 2>&1; tail -5 gpurun_out/rs_variants.json
timeout 60 python scripts/bench_train.py --math tf32 > gpurun_out/rs_train.json 2>&1; tail -2 gpurun_out/rs_train.json
timeout 60 python scripts/bench_int8.py 2>&1 | tail -1 | tee gpurun_out/rs_int8.json
timeout 60 python scripts/profile_int8.py 25 1 2>&1 | tail -20 | tee gpurun_out/rs_int8_profile.txt
