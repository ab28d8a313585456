#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; tail -2 gpurun_out/pytest_gpu.log
for V in "4 6" "3 8" "5 8"; do
  set -- $V
  echo "=== IAMRX_GS_MINB=$1 IAMRX_GSRB_MINB=$2"
  IAMRX_GS_MINB=$1 IAMRX_GSRB_MINB=$2 timeout 300 python scripts/kernel_table.py 256 3 2>&1 | head -12
done | tee gpurun_out/variants.txt
