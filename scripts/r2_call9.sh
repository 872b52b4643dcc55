#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-fused_v1}
stage() {  # name, timeout, command...
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout -k 10 "$to" "$@" > "gpurun_out/$name.log" 2>&1
  local rc=$?
  echo "$name rc=$rc" | tee -a gpurun_out/summary.txt
  tail -n 6 "gpurun_out/$name.log" | cut -c1-700
  return $rc
}
: > gpurun_out/summary.txt
stage ${tag}_tc 900 python -m pytest tests/test_gpu_b_tc.py -x -q -m gpu || { cat gpurun_out/summary.txt; exit 0; }
stage ${tag}_tests 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_b_tc.py
stage ${tag}_bench 900 python bench.py --steps 10 --warmup 3
stage ${tag}_bench_matrix 900 python bench.py --steps 10 --warmup 3 --rank-mode matrix --no-cpu-baseline --no-gpu-eager-baseline
cat gpurun_out/summary.txt
