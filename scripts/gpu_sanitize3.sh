#!/bin/bash
# compute-sanitizer over the fused i2t kernel (scan_i2t_tc2.cu) at fixture sizes: memcheck and racecheck (shared-memory
# scratches shared through warp barriers, metadata ring, cluster barriers), and memcheck of the streamed image uploads.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # name tool tests...
  local name=$1 tool=$2; shift 2
  timeout -k 10 900 compute-sanitizer --tool $tool --error-exitcode 1 --print-limit 20 python -m pytest "$@" > gpurun_out/sanitizer3_$name.log 2>&1
  echo "$name ($tool) rc=$?"; tail -4 gpurun_out/sanitizer3_$name.log | cut -c1-200
}
run memcheck_i2t memcheck tests/test_gpu_b_tc.py -x -q -m gpu -k "fused_i2t"
run racecheck_i2t racecheck tests/test_gpu_b_tc.py -x -q -m gpu -k "fused_i2t_matches_two_phase_and_oracle and Mean"
run memcheck_stream memcheck tests/test_gpu_b_tc.py -x -q -m gpu -k "streamed"
