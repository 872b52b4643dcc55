#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
stage() {  # name, timeout, command...
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout -k 10 "$to" "$@" > "gpurun_out/$name.log" 2>&1
  local rc=$?
  echo "$name rc=$rc" | tee -a gpurun_out/summary.txt
  tail -n 8 "gpurun_out/$name.log"
  return $rc
}
: > gpurun_out/summary.txt
stage r2c4_tc 600 python -m pytest tests/test_gpu_b_tc.py tests/test_gpu_c_fullsize.py -x -q -m gpu || { cat gpurun_out/summary.txt; exit 0; }
stage r2c4_roles 300 python scripts/role_profile2.py 1000 5000
stage r2c4_bench 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline
cat gpurun_out/summary.txt
