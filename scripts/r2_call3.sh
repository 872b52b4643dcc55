#!/bin/bash
# Round-2 GPU call 3: bring-up of the CTA-pair (cta_group::2) score kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
stage() {  # name, timeout, command...
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout -k 10 "$to" "$@" > "gpurun_out/$name.log" 2>&1
  local rc=$?
  echo "$name rc=$rc" | tee -a gpurun_out/summary.txt
  tail -n 15 "gpurun_out/$name.log"
  return $rc
}
: > gpurun_out/summary.txt
stage r2c3_smoke 300 python __graft_entry__.py smoke || { cat gpurun_out/summary.txt; exit 0; }
stage r2c3_tc 600 python -m pytest tests/test_gpu_b_tc.py -x -q -m gpu || { cat gpurun_out/summary.txt; exit 0; }
stage r2c3_tests 1200 python -m pytest tests -q -m gpu --deselect tests/test_gpu_b_tc.py
stage r2c3_roles 300 python scripts/role_profile2.py 1000 5000
stage r2c3_roles_v8 300 python scripts/role_profile.py 1000 5000 1
stage r2c3_bench 900 python bench.py --steps 5 --warmup 3
ITR_B200_SCORE_KERNEL=single stage r2c3_bench_single 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline
cat gpurun_out/summary.txt
