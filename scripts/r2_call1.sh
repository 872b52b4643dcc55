#!/bin/bash
# Round-2 GPU call 1: MMA cost model (1-CTA and CTA-pair), the whole GPU suite, bench with the new legs.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
stage() {  # name, timeout, command...
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout -k 10 "$to" "$@" > "gpurun_out/$name.log" 2>&1
  local rc=$?
  echo "$name rc=$rc" | tee -a gpurun_out/summary.txt
  tail -n 12 "gpurun_out/$name.log"
}
: > gpurun_out/summary.txt
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
stage r2_mma2_microbench 120 python scripts/mma2_microbench.py
stage r2_tests 1500 python -m pytest tests -m gpu -x -q
stage r2_bench 900 python bench.py --steps 5 --warmup 3
stage r2_bench_ref 600 python bench.py --impl reference --steps 2 --warmup 1
cat gpurun_out/summary.txt
