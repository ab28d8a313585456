#!/bin/bash
# one GPU-box visit: gpu tests, smoke, kernel table, bench (both arms), ncu launch list of the timed step.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 300 python scripts/kernel_table.py 256 3 > gpurun_out/kernel_table.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
if [ "$1" != "quick" ]; then
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
fi
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; cat gpurun_out/kernel_table.txt; cat gpurun_out/bench.json; wc -l gpurun_out/launches.csv
