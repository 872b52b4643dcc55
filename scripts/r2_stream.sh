#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_b_tc.py tests/test_gpu_e_dropin.py -x -q -m gpu -k "streamed or row_ranges or dropin or fused" > gpurun_out/stream_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/stream_tests.log | cut -c1-300
timeout -k 10 900 python bench.py > gpurun_out/stream_bench.json 2> gpurun_out/stream_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
j=json.loads(open('gpurun_out/stream_bench.json').read().strip().split('\n')[-1])
print("device ms", j['ms_per_step'], "e2e ms", j['e2e']['ms_per_step'], "dropin", j.get('e2e_dropin',{}).get('ms_per_step'), "frac", j['roofline']['frac'], "clocks", j['clocks']['sm_mhz'])
PY
