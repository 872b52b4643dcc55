#!/bin/bash
# Staged GPU bring-up: every stage is its own process under its own timeout, so a trap or a
# hang in one stage neither hides the others nor wedges the box.  Logs go to gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
stage() {  # name, timeout, command...
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout -k 10 "$to" "$@" > "gpurun_out/$name.log" 2>&1
  local rc=$?
  echo "$name rc=$rc" | tee -a gpurun_out/summary.txt
  tail -n 25 "gpurun_out/$name.log"
}
: > gpurun_out/summary.txt
stage simt 900 python -m pytest tests/test_gpu_a_simt.py -x -q -m gpu
stage tc_affinity 300 python -m pytest tests/test_gpu_b_tc.py -x -q -m gpu -k affinity
stage tc_rest 900 python -m pytest tests/test_gpu_b_tc.py -q -m gpu -k "not affinity"
stage fullsize 900 python -m pytest tests/test_gpu_c_fullsize.py -q -m gpu
stage smoke 300 python __graft_entry__.py smoke
stage bench_small 600 python bench.py --n-img 1000 --n-cap 5000 --steps 3 --warmup 2 --no-cpu-baseline
stage bench_full 1200 python bench.py --steps 3 --warmup 3
cat gpurun_out/summary.txt
