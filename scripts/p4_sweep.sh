#!/bin/bash
# ablation / profile sweep of the planar-C4 kernel on the layer-1 shape (first 2 lines of bench_p4)
echo "== default"; timeout 200 python scripts/bench_p4.py 10 2>&1 | grep "^p4"
echo "== prof"; QBN_P4_PROF=1 timeout 200 python scripts/bench_p4.py 10 2>&1 | grep "p4 prof" | awk '{k=$3" "$4" "$5" "$6; if(!(k in s)){s[k]=1; print}}'
for d in 1 2 4 8 3 7 15; do echo "== DBG $d"; QBN_P4_DBG=$d timeout 200 python scripts/bench_p4.py 10 2>&1 | grep "^p4" | head -7; done
