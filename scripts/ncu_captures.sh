#!/bin/bash
# The three ncu captures behind profiles/r02_p4_dram_traffic.json, r02_lrt_kernels_ncu_full.csv and r02_launches.csv (one GPU):
#   gpurun --timeout 2400 -- bash scripts/ncu_captures.sh
mkdir -p gpurun_out
# (1) DRAM traffic of the benchmarked configuration, final kernels
timeout 600 ncu --set full --clock-control none -k regex:umma_conv -c 44 -o /tmp/r02_p4 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-train --no-int8 --no-gpu-eager > gpurun_out/c13_ncu_eval.log 2>&1
ncu -i /tmp/r02_p4.ncu-rep --page raw --csv > gpurun_out/r02_p4_raw.csv 2>/dev/null
# (2) the LRT training kernels of one eager B=256 step (4th step of profile_train.py: 59 matching launches per step)
timeout 600 ncu --set full --clock-control none -k regex:"umma_conv_p4_kernel|umma_wgrad_p4" -s 177 -c 59 -o /tmp/r02_lrt python scripts/profile_train.py tf32 > gpurun_out/c13_ncu_lrt.log 2>&1
ncu -i /tmp/r02_lrt.ncu-rep --page raw --csv > gpurun_out/r02_lrt_raw.csv 2>/dev/null
# (3) launch list of the bench command
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 700 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-train --no-int8 > gpurun_out/c13_launches.log 2>&1
ls -la gpurun_out/r02_*.csv
