#!/bin/bash
# deep-ghost red-black sweeps of the cell-centred multigrid: parity tests, 8-box 256^3 on one GPU off / on
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_solvers.py tests/test_step.py tests/test_forcing.py -m gpu -x -q 2>&1 | tail -3
for D in 0 1; do
  IAMRX_CELL_DEEP=$D timeout 600 python scripts/multibox_bench.py 256 2 2 2 5 > gpurun_out/multibox_celldeep$D.txt 2>&1; echo "CELL_DEEP=$D"; head -8 gpurun_out/multibox_celldeep$D.txt; tail -1 gpurun_out/multibox_celldeep$D.txt
done
