#!/bin/bash
# Final ncu evidence of the round: (1) launch list of the default bench command with the final code, (2) one --set full
# capture of the fused i2t kernel on one COCO fold.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
  --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-eager-baseline \
  > gpurun_out/launches_final.log 2>&1
echo "launch list rc=$?"
bash scripts/r2_i2t_prof.sh
