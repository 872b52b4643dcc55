#!/bin/bash
# Round-2 ncu evidence for the CTA-pair score kernel, on the HEADLINE shape (5000 x 25000):
#  (1) launch list (device time per launch) of the default bench command,
#  (2) one `--set full` capture of scan_t2i_tc2_kernel (ncu replays the kernel ~40x: ~6 s of kernel time),
#      exported to raw / source CSV on the box (the .ncu-rep itself stays there).
# Usage: scripts/gpu_profile2.sh <tag>
cd "$(dirname "$0")/.."
tag=${1:-pair}
mkdir -p gpurun_out
timeout -k 10 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
  --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-eager-baseline \
  > gpurun_out/launches_$tag.log 2>&1
echo "launch list rc=$?"
timeout -k 10 1500 ncu --set full --clock-control none --import-source on -k regex:scan_t2i_tc2_kernel -s 1 -c 1 \
  -o gpurun_out/prof_$tag -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-eager-baseline \
  > gpurun_out/prof_$tag.log 2>&1
echo "full capture rc=$?"
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$tag.ncu-rep --page source --csv > gpurun_out/prof_${tag}_src.csv 2>/dev/null
rm -f gpurun_out/prof_$tag.ncu-rep
ls -la gpurun_out | tail -5
