#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "" "ITR_B200_I2T_SKIP=1" "ITR_B200_I2T_BAND=8" "ITR_B200_I2T_BAND=16" "ITR_B200_I2T_BAND=64" "ITR_B200_I2T_BAND=8 ITR_B200_I2T_SKIP=1"; do
  echo "== $v"; env $v timeout 120 python scripts/i2t_time.py 2>&1 | head -1
done
