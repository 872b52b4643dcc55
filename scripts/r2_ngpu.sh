#!/bin/bash
# multi-GPU bench: bash scripts/r2_ngpu.sh N [tag]
cd "$(dirname "$0")/.."
N=${1:-2}; tag=${2:-r2}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_b_tc.py -x -q -m gpu -k "row_ranges or fused" > gpurun_out/ngpu_${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/ngpu_${tag}_tests.log
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu_${tag}.log 2>&1; echo "bench rc=$?"
tail -c 2500 gpurun_out/bench_${N}gpu_${tag}.log
ITR_B200_GATHER=nccl timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu_${tag}_nccl.log 2>&1; echo "bench nccl rc=$?"
tail -c 600 gpurun_out/bench_${N}gpu_${tag}_nccl.log
