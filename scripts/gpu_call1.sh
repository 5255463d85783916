#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lrt_p4.py -q -x 2>&1 | tail -40 > gpurun_out/c1_lrt_tests.log
timeout 300 python -m pytest tests/test_gpu_lrt_p4.py -q 2>&1 | tail -60 > gpurun_out/c1_lrt_tests_all.log
timeout 120 python scripts/profile_train.py tf32 > gpurun_out/c1_profile_train.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager --pdl 0 > gpurun_out/c1_bench_pdl0.json 2> gpurun_out/c1_bench_pdl0.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-train --pdl 1 > gpurun_out/c1_bench_pdl1.json 2> gpurun_out/c1_bench_pdl1.err
timeout 600 ncu --set full --clock-control none -k regex:umma_conv -c 44 -o /tmp/r02_p4 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-train --no-int8 --no-gpu-eager > gpurun_out/c1_ncu.log 2>&1
ncu -i /tmp/r02_p4.ncu-rep --page raw --csv > gpurun_out/r02_p4_raw.csv 2>/dev/null
ls -la /tmp/r02_p4.ncu-rep >> gpurun_out/c1_ncu.log
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/c1_suite.log
tail -5 gpurun_out/c1_lrt_tests.log
# role-cycle accounting of the planar kernel at 50-sample chunks (tuning build, scratch copy only)
timeout 120 python scripts/bench_p4.py 50 2>&1 | grep "^p4" > gpurun_out/c1_p4_layers.txt
QBN_TUNING=1 python -c "from qbn_b200 import _build; _build.build_lib(force=True)" > /dev/null 2>&1
QBN_P4_PROF=1 timeout 200 python scripts/bench_p4.py 50 2>&1 | grep "p4 prof" | awk '{k=$3" "$4" "$5" "$6" "$7; if(!(k in s)){s[k]=1; print}}' > gpurun_out/c1_p4_prof.txt
