#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_b_tc.py -x -q -m gpu -k "fused_i2t or edge_shapes or generic" > gpurun_out/i2t_tests.log 2>&1; echo "i2t tests rc=$?"; tail -25 gpurun_out/i2t_tests.log | cut -c1-300
timeout -k 10 600 python scripts/i2t_time.py > gpurun_out/i2t_time.log 2>&1; cat gpurun_out/i2t_time.log
timeout -k 10 600 python bench.py --config 4 > gpurun_out/i2t_cfg4.log 2>&1; echo "cfg4 rc=$?"; tail -2 gpurun_out/i2t_cfg4.log | cut -c1-700

