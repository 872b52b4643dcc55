#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for k in 1 2 3; do
timeout -k 10 200 python -m pytest tests/test_gpu_b_tc.py -x -q -m gpu -k "fused_i2t" > gpurun_out/i2t_tests$k.log 2>&1; echo "i2t tests $k rc=$?"; tail -1 gpurun_out/i2t_tests$k.log | cut -c1-200
timeout -k 10 120 python scripts/i2t_time.py > gpurun_out/i2t_time$k.log 2>&1; echo "time $k rc=$?"; head -1 gpurun_out/i2t_time$k.log | cut -c1-100
done
