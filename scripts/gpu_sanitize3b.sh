#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 compute-sanitizer --tool racecheck --error-exitcode 1 --print-limit 20 python -m pytest tests/test_gpu_b_tc.py -x -q -m gpu -k "fused_i2t_matches_two_phase_and_oracle and Mean" > gpurun_out/sanitizer3_racecheck_i2t.log 2>&1
echo "racecheck rc=$?"; grep "=========" gpurun_out/sanitizer3_racecheck_i2t.log | grep -v "     at\|     in\|Host Frame" | head -12
bash scripts/r2_i2t_probe.sh
