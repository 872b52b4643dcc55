#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 1800 python -m pytest tests -m gpu -q > gpurun_out/r2_tests.log 2>&1
echo "tests rc=$?"; tail -n 30 gpurun_out/r2_tests.log
