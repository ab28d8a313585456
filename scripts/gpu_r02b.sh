#!/bin/bash
# round-2 GPU visit B: boundary-condition tests on the device + RayleighTaylor single-level timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_bc.py tests/test_kernels.py tests/test_solvers.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_bc.log
tail -5 gpurun_out/pytest_gpu_bc.log
timeout 600 python scripts/rt_bench.py > gpurun_out/rt_bench.txt 2>&1; tail -30 gpurun_out/rt_bench.txt
