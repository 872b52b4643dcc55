#!/bin/bash
# ncu evidence: (1) launch list (device time per launch) of the default bench command, (2) one full capture of the
# tcgen05 score kernel on a smaller shape (ncu replays the kernel ~40x).  Usage: scripts/gpu_profile.sh <tag>
cd "$(dirname "$0")/.."
tag=${1:-prof}
mkdir -p gpurun_out
timeout -k 10 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
  --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline \
  > gpurun_out/launches_$tag.log 2>&1
echo "launch list rc=$?"
timeout -k 10 1200 ncu --set full --clock-control none --import-source on -k regex:scan_t2i_tc_kernel -s 1 -c 1 \
  -o gpurun_out/prof_$tag -f python bench.py --n-img 1000 --n-cap 5000 --steps 1 --warmup 1 --no-cpu-baseline \
  > gpurun_out/prof_$tag.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -6
