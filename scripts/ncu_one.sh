#!/bin/bash
# ncu --set full capture of one kernel family launched in isolation: scripts/ncu_one.sh <kernel regex> <kb_one.py target> [launch skip]
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 -s ${3:-2} -c 2 -f -o gpurun_out/prof_$2 \
   python scripts/kb_one.py $2 256 3 > gpurun_out/ncu_$2.log 2>&1
tail -3 gpurun_out/ncu_$2.log; ls -la gpurun_out/prof_$2.ncu-rep
