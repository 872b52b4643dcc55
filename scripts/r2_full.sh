#!/bin/bash
# full validation pass on one B200: GPU tests, smoke, headline bench, config 4
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -x -q -m gpu > gpurun_out/full_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/full_tests.log | cut -c1-300
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/full_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/full_smoke.log | cut -c1-300
timeout -k 10 900 python bench.py > gpurun_out/full_bench.json 2> gpurun_out/full_bench.err; echo "bench rc=$?"; tail -1 gpurun_out/full_bench.json | cut -c1-1500
timeout -k 10 600 python bench.py --config 4 > gpurun_out/full_cfg4.json 2> gpurun_out/full_cfg4.err; echo "cfg4 rc=$?"; tail -1 gpurun_out/full_cfg4.json | cut -c1-900
