#!/bin/bash
# short GPU-box visit: gpu tests, kernel micro-benchmarks, per-kernel table of the step
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
timeout 300 python scripts/kbench.py 256 20 > gpurun_out/kbench.txt 2>&1
if [ "$1" != "nokt" ]; then timeout 300 python scripts/kernel_table.py 256 3 > gpurun_out/kernel_table.txt 2>&1; fi
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/kbench.txt; if [ "$1" != "nokt" ]; then head -24 gpurun_out/kernel_table.txt; fi
