#!/bin/bash
# round-2 GPU visit A: gpu tests, smoke, fp64 peak, kernel table, bench (both arms), ncu launch list, full captures of the shipping kernels
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt; free -g | head -2 >> gpurun_out/nproc.txt
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; cat gpurun_out/smoke.log | tail -2
timeout 120 python scripts/fp64_peak.py gpurun_out/fp64_peak.json
timeout 300 python scripts/kernel_table.py 256 3 > gpurun_out/kernel_table.txt 2>&1; head -30 gpurun_out/kernel_table.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --e2e-steps 0 --prof-steps 0 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
wc -l gpurun_out/launches.csv
for K in aofs_tile_kernel adotx_march_kernel ev_edge_kernel ev_corner_kernel ev_final_kernel apply2_kernel tensor_cross_kernel gs_sweep_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off \
     -k regex:$K -c 1 -f -o gpurun_out/prof_$K python bench.py --steps 1 --warmup 0 --e2e-steps 0 --prof-steps 0 --no-cpu-baseline \
     > gpurun_out/ncu_$K.log 2>&1
  tail -1 gpurun_out/ncu_$K.log
done
timeout 600 python scripts/kbench.py --spec > gpurun_out/kbench_spec.txt 2>&1; tail -20 gpurun_out/kbench_spec.txt
ls -la gpurun_out/
