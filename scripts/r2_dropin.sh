#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_e_dropin.py tests/test_capi_symbols.py -x -q -m "gpu or not gpu" > gpurun_out/dropin_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/dropin_tests.log | cut -c1-300
timeout -k 10 900 python bench.py --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/dropin_bench.json 2> gpurun_out/dropin_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
j=json.loads(open('gpurun_out/dropin_bench.json').read().strip().split('\n')[-1])
print("device ms", j['ms_per_step'], "e2e ms", j['e2e']['ms_per_step'], "dropin", j.get('e2e_dropin'))
PY
