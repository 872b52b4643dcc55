#!/bin/bash
# one short, tightly bounded run of the fused i2t kernel before anything longer is spent on it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 60 python scripts/i2t_time.py > gpurun_out/i2t_probe.log 2>&1; rc=$?; echo "probe rc=$rc"; head -3 gpurun_out/i2t_probe.log | cut -c1-200
[ $rc -ne 0 ] && exit 1
timeout -k 5 90 python -m pytest tests/test_gpu_b_tc.py -x -q -m gpu -k "fused_i2t" > gpurun_out/i2t_probe_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/i2t_probe_tests.log | cut -c1-300
timeout -k 5 60 python scripts/i2t_flaky.py 2>&1 | tail -6
