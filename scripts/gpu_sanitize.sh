#!/bin/bash
# compute-sanitizer over the kernels at golden-fixture sizes (memcheck, then racecheck on the tcgen05 kernel).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 compute-sanitizer --tool memcheck --error-exitcode 1 --print-limit 20 \
  python -m pytest tests/test_gpu_b_tc.py tests/test_gpu_a_simt.py -x -q -m gpu -k "golden or hinge or ranking or ragged" \
  > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -6 gpurun_out/sanitizer_memcheck.log
timeout -k 10 900 compute-sanitizer --tool racecheck --error-exitcode 1 --print-limit 20 \
  python -m pytest tests/test_gpu_b_tc.py -x -q -m gpu -k "tc_golden" \
  > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -6 gpurun_out/sanitizer_racecheck.log
