#!/bin/bash
# GPU tests, continuing past failures, bounded
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -q -m gpu > gpurun_out/full_tests.log 2>&1; echo "gpu tests rc=$?"; tail -6 gpurun_out/full_tests.log | cut -c1-300
