#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-pair_v5}
stage() {  # name, timeout, command...
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout -k 10 "$to" "$@" > "gpurun_out/$name.log" 2>&1
  local rc=$?
  echo "$name rc=$rc" | tee -a gpurun_out/summary.txt
  tail -n 4 "gpurun_out/$name.log" | cut -c1-600
  return $rc
}
: > gpurun_out/summary.txt
stage ${tag}_tc 600 python -m pytest tests/test_gpu_b_tc.py tests/test_gpu_c_fullsize.py -x -q -m gpu || { cat gpurun_out/summary.txt; exit 0; }
stage ${tag}_roles 300 python scripts/role_profile2.py 1000 5000
stage ${tag}_bench1 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline
stage ${tag}_bench2 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline
ITR_B200_SCORE_KERNEL=single stage ${tag}_bench_single 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline
cat gpurun_out/summary.txt
