#!/bin/bash
# compute-sanitizer memcheck over a few representative GPU tests (planar eval engine with unit windows, planar LRT forward/backward,
# int8 planar engine).  Slow (10-50x): each pytest selection runs under its own timeout.   gpurun --timeout 900 -- bash scripts/sanitize.sh
mkdir -p gpurun_out
run() {
  name=$1; shift
  timeout 280 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/san_$name.log python -m pytest "$@" -x -q > gpurun_out/san_$name.out 2>&1
  echo "$name: rc=$? $(tail -1 gpurun_out/san_$name.out) | $(grep -c 'Invalid\|out of bounds\|misaligned' gpurun_out/san_$name.log) findings | $(tail -1 gpurun_out/san_$name.log)"
}
run window tests/test_gpu_models.py -k "unit_window and 8-6-True"
run lrt tests/test_gpu_lrt_p4.py -k "lrt_p4_backward or non_square or staging"
run p4 tests/test_gpu_p4.py -k "stride1 or shortcut or stacked"
run i8 tests/test_gpu_i8_full_resnet.py -k "planar_engine_equals"
