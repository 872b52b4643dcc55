#!/bin/bash
# ncu evidence for the bench command: (1) launch list with device times, (2) one full capture of the
# tcgen05 score kernel.  Usage: scripts/gpu_profile.sh <tag> [n_img n_cap]
cd "$(dirname "$0")/.."
tag=${1:-prof}; n_img=${2:-1000}; n_cap=${3:-5000}
mkdir -p gpurun_out
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches_$tag.csv python bench.py --n-img $n_img --n-cap $n_cap --steps 2 --warmup 1 --no-cpu-baseline \
  > gpurun_out/launches_$tag.log 2>&1
echo "launch list rc=$?"
timeout -k 10 1200 ncu --set full --clock-control none --import-source on -k regex:scan_t2i_tc_kernel -s 1 -c 1 \
  -o gpurun_out/prof_$tag -f python bench.py --n-img $n_img --n-cap $n_cap --steps 1 --warmup 1 --no-cpu-baseline \
  > gpurun_out/prof_$tag.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -8
