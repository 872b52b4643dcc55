#!/bin/bash
# compute-sanitizer over the round-2 kernels at fixture sizes: the CTA-pair score kernel in its three modes (memcheck +
# racecheck: shared-memory row walk, cluster barriers) and the cooperative VSE++ step (memcheck + racecheck).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # name tool tests...
  local name=$1 tool=$2; shift 2
  timeout -k 10 1200 compute-sanitizer --tool $tool --error-exitcode 1 --print-limit 20 python -m pytest "$@" > gpurun_out/sanitizer2_$name.log 2>&1
  echo "$name ($tool) rc=$?"; tail -5 gpurun_out/sanitizer2_$name.log
}
run memcheck_pair memcheck tests/test_gpu_b_tc.py -x -q -m gpu -k "tc_golden or fused_ranking or row_ranges or ragged"
run racecheck_pair racecheck tests/test_gpu_b_tc.py -x -q -m gpu -k "tc_golden or fused_ranking_long"
run memcheck_vse memcheck tests/test_gpu_a_simt.py -x -q -m gpu -k "fused_vse_step or hinge"
run racecheck_vse racecheck tests/test_gpu_a_simt.py -x -q -m gpu -k "fused_vse_step"
