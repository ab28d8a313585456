#!/bin/bash
# deep-ghost fused nodal sweep on block layouts: parity tests, then 256^3 as 2x2x2 boxes on one GPU with the path off / on
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_solvers.py tests/test_step.py tests/test_level.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_deep.log; cat gpurun_out/pytest_deep.log
for D in 0 1; do
  IAMRX_NODAL_DEEP=$D timeout 600 python scripts/multibox_bench.py 256 2 2 2 5 > gpurun_out/multibox_deep$D.txt 2>&1; echo "DEEP=$D"; cat gpurun_out/multibox_deep$D.txt | tail -18
done
IAMRX_NODAL_DEEP=1 timeout 600 python scripts/multibox_bench.py 256 4 4 4 5 > gpurun_out/multibox_deep1_444.txt 2>&1; head -1 gpurun_out/multibox_deep1_444.txt
IAMRX_NODAL_DEEP=0 timeout 600 python scripts/multibox_bench.py 256 4 4 4 5 > gpurun_out/multibox_deep0_444.txt 2>&1; head -1 gpurun_out/multibox_deep0_444.txt
