#!/bin/bash
# advection-focused GPU visit: advection parity tests + micro-benchmarks
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels.py -m gpu -x -q -k "aofs or extrap" 2>&1 | tail -15 > gpurun_out/pytest_adv.log
timeout 300 python scripts/kbench.py 256 10 > gpurun_out/kbench.txt 2>&1
tail -6 gpurun_out/pytest_adv.log; cat gpurun_out/kbench.txt
