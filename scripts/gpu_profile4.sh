#!/bin/bash
# --set full capture of the final pair kernel (counting mode) on the headline shape (5000 x 25000)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 1500 ncu --set full --clock-control none --import-source on -k regex:scan_t2i_tc2_kernel -s 1 -c 1 \
  -o gpurun_out/prof_pair_final -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-eager-baseline \
  > gpurun_out/prof_pair_final.log 2>&1
echo "full capture rc=$?"
ncu -i gpurun_out/prof_pair_final.ncu-rep --page raw --csv > gpurun_out/prof_pair_final_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_pair_final.ncu-rep --page source --csv > gpurun_out/prof_pair_final_src.csv 2>/dev/null
rm -f gpurun_out/prof_pair_final.ncu-rep
