#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/profile_int8.py 50 1 planar > gpurun_out/c11_int8_profile.txt 2>&1
QBN_TUNING=1 python -c "from qbn_b200 import _build; _build.build_lib(force=True)" > /dev/null 2>&1
QBN_TUNING=1 QBN_P4_PROF=1 timeout 200 python scripts/profile_int8.py 50 1 planar 2>&1 | grep "p4 prof" | awk '{k=$4" "$5" "$6" "$7; if(!(k in s)){s[k]=1; print}}' > gpurun_out/c11_int8_prof_roles.txt
tail -25 gpurun_out/c11_int8_profile.txt; cat gpurun_out/c11_int8_prof_roles.txt | cut -c1-400
