#!/bin/bash
# steady-state (second step: the density is no longer bitwise constant) launch list and GSRB capture; full GPU suite
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/z2_pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv --log-file gpurun_out/z2_launches.csv \
   python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --prof-steps 0 > gpurun_out/z2_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gsrb_kernel -s 1 -c 1 -f -o gpurun_out/z2_prof_gsrb \
   python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --prof-steps 0 > gpurun_out/z2_ncu_gsrb.log 2>&1
python scripts/ncu_summary.py gpurun_out/z2_prof_gsrb.ncu-rep > gpurun_out/z2_ncu_full_gsrb_kernel.txt 2>/dev/null
head -4 gpurun_out/z2_ncu_full_gsrb_kernel.txt
tail -3 gpurun_out/z2_pytest_gpu.log
