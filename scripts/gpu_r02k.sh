#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/dsl2d_diag.py 512 1 > gpurun_out/dsl512.log 2>&1; tail -8 gpurun_out/dsl512.log
timeout 300 python scripts/dsl2d_diag.py 1024 2 > gpurun_out/dsl1024.log 2>&1; grep -v " iter " gpurun_out/dsl1024.log | tail -8; grep " iter " gpurun_out/dsl1024.log | head -14; grep " iter " gpurun_out/dsl1024.log | tail -4
