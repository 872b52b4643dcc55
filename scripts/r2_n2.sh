#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu_final.log 2>&1; echo "bench rc=$?"
python - "$N" <<'PY'
import json,sys
n=sys.argv[1]
lines=[l for l in open('gpurun_out/bench_%sgpu_final.log'%n).read().split('\n') if l.startswith('{')]
j=json.loads(lines[-1])
print("N",n,"device ms", j['ms_per_step'], "e2e ms", j['e2e']['ms_per_step'], "frac", j['roofline']['frac'], "clocks", j['clocks']['sm_mhz'], "recall", j['recall_check'], "identical", j['rank_mode']['identical_ranks'])
print(j['device_breakdown_ms'])
PY
