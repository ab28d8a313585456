#!/bin/bash
# last pass of round 2: GPU suite, bench lines (TaylorGreen with CPU baseline, RayleighTaylor, HIT, DoubleShearLayer 2-D), kernel table
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/z4_pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/z4_bench.json 2> gpurun_out/z4_bench.err
python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline --kernel-table > /dev/null 2> gpurun_out/z4_kernel_table.err
grep -A60 "^# kernel table" gpurun_out/z4_kernel_table.err > gpurun_out/z4_kernel_table.txt
python bench.py --problem rt --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/z4_bench_rt.json 2> gpurun_out/z4_bench_rt.err
python bench.py --problem hit --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/z4_bench_hit.json 2> gpurun_out/z4_bench_hit.err
python bench.py --problem dsl2d --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/z4_bench_dsl2d.json 2> gpurun_out/z4_bench_dsl2d.err
python - <<'PY'
import json
for f in ("", "_rt", "_hit", "_dsl2d"):
    try:
        d = json.loads(open("gpurun_out/z4_bench%s.json" % f).read().strip().splitlines()[-1])
        print(f or "tg", round(d["ms_per_step"], 2), round(d["value"] / 1e6, 1), "e2e", round((d["e2e"]["value"] or 0) / 1e6, 1), round(d["roofline"]["frac"], 3), d["config"]["mg_iters_last_step"], d.get("cpu_baseline", {}).get("value"), d.get("clocks"))
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/z4_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
