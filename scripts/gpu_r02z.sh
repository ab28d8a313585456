#!/bin/bash
# Final round-2 measurement pass: GPU suite, bench (TaylorGreen 256^3) with CPU baseline, kernel table, ncu launch list, ncu --set full
# captures of the kernels changed in the last part of the round.  Outputs in gpurun_out/z_*.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/z_pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err
python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline --kernel-table > /dev/null 2> gpurun_out/z_kernel_table.err
grep -A60 "^# kernel table" gpurun_out/z_kernel_table.err > gpurun_out/z_kernel_table.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv --log-file gpurun_out/z_launches.csv \
   python bench.py --steps 1 --warmup 0 --e2e-steps 0 --no-cpu-baseline --prof-steps 0 > gpurun_out/z_launches.log 2>&1
cap() {  # name regex skip
  timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$2 -s $3 -c 1 -f -o gpurun_out/z_prof_$1 \
     python bench.py --steps 1 --warmup 0 --e2e-steps 0 --no-cpu-baseline --prof-steps 0 > gpurun_out/z_ncu_$1.log 2>&1
  python scripts/ncu_summary.py gpurun_out/z_prof_$1.ncu-rep > gpurun_out/z_ncu_full_$1.txt 2>/dev/null
  head -3 gpurun_out/z_ncu_full_$1.txt
}
cap gsrb_kernel gsrb_kernel 1
cap gs_sweep_kernel gs_sweep_kernel 0
cap adotx_march_kernel adotx_march_kernel 1
cap apply2_kernel apply2_kernel 2
cap nd_interp8_kernel nd_interp8_kernel 0
python bench.py --problem rt --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/z_bench_rt.json 2> gpurun_out/z_bench_rt.err
python bench.py --problem hit --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/z_bench_hit.json 2> gpurun_out/z_bench_hit.err
python bench.py --problem dsl2d --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/z_bench_dsl2d.json 2> gpurun_out/z_bench_dsl2d.err
python - <<'PY'
import json
for f in ("", "_rt", "_hit", "_dsl2d"):
    try:
        d = json.loads(open("gpurun_out/z_bench%s.json" % f).read().strip().splitlines()[-1])
        print(f or "tg", round(d["ms_per_step"], 2), round(d["value"] / 1e6, 1), "e2e", round((d["e2e"]["value"] or 0) / 1e6, 1), round(d["roofline"]["frac"], 3), d["config"]["mg_iters_last_step"], d.get("cpu_baseline", {}).get("value"))
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/z_pytest_gpu.log
