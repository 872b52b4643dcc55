#!/bin/bash
# pair kernel with the shared-memory segmented sums: parity, role counters, bench, one full ncu capture (1000 x 5000)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-pair_v3}
stage() {  # name, timeout, command...
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout -k 10 "$to" "$@" > "gpurun_out/$name.log" 2>&1
  local rc=$?
  echo "$name rc=$rc" | tee -a gpurun_out/summary.txt
  tail -n 8 "gpurun_out/$name.log"
  return $rc
}
: > gpurun_out/summary.txt
stage ${tag}_tc 600 python -m pytest tests/test_gpu_b_tc.py tests/test_gpu_c_fullsize.py -x -q -m gpu || { cat gpurun_out/summary.txt; exit 0; }
stage ${tag}_roles 300 python scripts/role_profile2.py 1000 5000
stage ${tag}_bench 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline
stage ${tag}_ncu 900 ncu --set full --clock-control none --import-source on -k regex:scan_t2i_tc2_kernel -s 1 -c 1 \
  -o gpurun_out/prof_$tag -f python bench.py --n-img 1000 --n-cap 5000 --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-eager-baseline
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$tag.ncu-rep --page source --csv > gpurun_out/prof_${tag}_src.csv 2>/dev/null
rm -f gpurun_out/prof_$tag.ncu-rep
cat gpurun_out/summary.txt
