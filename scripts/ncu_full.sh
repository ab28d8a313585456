#!/bin/bash
# ncu --set full captures of single finest-level launches inside the timed step of bench.py
mkdir -p gpurun_out
for K in "$@"; do
  timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off \
     -k regex:$K -c 1 -f -o gpurun_out/prof_$K python bench.py --steps 1 --warmup 0 --e2e-steps 0 --no-cpu-baseline \
     > gpurun_out/ncu_$K.log 2>&1
  tail -2 gpurun_out/ncu_$K.log
done
ls -la gpurun_out/*.ncu-rep
