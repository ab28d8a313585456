#!/bin/bash
# multigrid-focused GPU visit: kernel + solver + step parity tests, micro-benchmarks, kernel table
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
timeout 300 python scripts/kbench.py 256 10 > gpurun_out/kbench.txt 2>&1
timeout 300 python scripts/kernel_table.py 256 3 > gpurun_out/kernel_table.txt 2>&1
tail -6 gpurun_out/pytest_gpu.log; cat gpurun_out/kbench.txt; head -16 gpurun_out/kernel_table.txt
