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
# SCAN training backward (coefficient kernel, GEMMs, block-diagonal term): memcheck + racecheck (shared-memory phases)
timeout -k 10 900 compute-sanitizer --tool memcheck --error-exitcode 1 --print-limit 20 \
  python -m pytest tests/test_gpu_d_backward.py -x -q -m gpu -k "golden or seeded or errors" \
  > gpurun_out/sanitizer_memcheck_backward.log 2>&1
echo "memcheck(backward) rc=$?"; tail -4 gpurun_out/sanitizer_memcheck_backward.log
timeout -k 10 900 compute-sanitizer --tool racecheck --error-exitcode 1 --print-limit 20 \
  python -m pytest tests/test_gpu_d_backward.py -x -q -m gpu -k "contrastive_loss" \
  > gpurun_out/sanitizer_racecheck_backward.log 2>&1
echo "racecheck(backward) rc=$?"; tail -4 gpurun_out/sanitizer_racecheck_backward.log
