#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
stage() {  # name, timeout, command...
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout -k 10 "$to" "$@" > "gpurun_out/$name.log" 2>&1
  local rc=$?
  echo "$name rc=$rc" | tee -a gpurun_out/summary.txt
  tail -n 5 "gpurun_out/$name.log" | cut -c1-900
  return $rc
}
: > gpurun_out/summary.txt
stage vse_tests 900 python -m pytest tests/test_gpu_a_simt.py tests/test_gpu_c_fullsize.py -x -q -m gpu





cat gpurun_out/summary.txt
