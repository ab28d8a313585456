#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/hit_diag.py 256 0.5 1e-10 1 > gpurun_out/hit_diag_vd05_tol10.log 2>&1; tail -25 gpurun_out/hit_diag_vd05_tol10.log
