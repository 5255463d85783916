#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_lrt_p4.py -q 2>&1 | tail -150 > gpurun_out/c2_lrt_tests_all.log
timeout 120 python scripts/dbg/wgrad_diag.py > gpurun_out/c2_wgrad_diag.txt 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager --pdl 0 > gpurun_out/c2_bench_pdl0.json 2> gpurun_out/c2_bench_pdl0.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-train --pdl 1 > gpurun_out/c2_bench_pdl1.json 2> gpurun_out/c2_bench_pdl1.err
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/c2_suite.log
timeout 120 python scripts/bench_p4.py 50 2>&1 | grep "^p4" > gpurun_out/c2_p4_layers.txt
# tuning build (scratch copy only): descriptor variants of the weight-gradient kernel, role-cycle accounting of the planar kernel
QBN_TUNING=1 python -c "from qbn_b200 import _build; _build.build_lib(force=True)" > /dev/null 2>&1
for v in 1 2 3 4 7; do QBN_TUNING=1 QBN_WG_V=$v timeout 60 python scripts/dbg/wgrad_diag.py >> gpurun_out/c2_wgrad_diag.txt 2>&1; done
QBN_TUNING=1 QBN_P4_PROF=1 timeout 200 python scripts/bench_p4.py 50 2>&1 | grep "p4 prof" | awk '{k=$3" "$4" "$5" "$6" "$7; if(!(k in s)){s[k]=1; print}}' > gpurun_out/c2_p4_prof.txt
QBN_TUNING=1 QBN_P4_RR=0 timeout 120 python scripts/bench_p4.py 50 2>&1 | grep "^p4" > gpurun_out/c2_p4_layers_rr0.txt
tail -5 gpurun_out/c2_lrt_tests_all.log
