#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python scripts/kernel_table.py 256 3 > gpurun_out/kernel_table.txt 2>&1; head -24 gpurun_out/kernel_table.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "
import json; b=json.load(open('gpurun_out/bench.json')); print(b['ms_per_step'], b['value'], b['e2e']['value'], b['roofline']['frac'], {k:(round(v['frac'],3), round(v['ms_per_step'],2)) for k,v in b['roofline']['other_kernels'].items()})"
timeout 600 python scripts/rt_bench.py > gpurun_out/rt_bench.txt 2>&1; head -16 gpurun_out/rt_bench.txt
