#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:scan_i2t_tc2_kernel -s 1 -c 1 \
  -o gpurun_out/prof_i2t -f python scripts/i2t_time.py > gpurun_out/prof_i2t.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/prof_i2t.ncu-rep --page raw --csv > gpurun_out/prof_i2t_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_i2t.ncu-rep --page source --csv > gpurun_out/prof_i2t_src.csv 2>/dev/null
rm -f gpurun_out/prof_i2t.ncu-rep
