#!/bin/bash
# 1-GPU visit: the full GPU test-suite, then the new bench lines (2-D DSL, forced HIT 256^3) and the default line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu_v2.log; cat gpurun_out/pytest_gpu_v2.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --problem dsl2d > gpurun_out/y_dsl2d.json 2> gpurun_out/y_dsl2d.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --problem hit > gpurun_out/y_hit.json 2> gpurun_out/y_hit.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --problem hit --no-forcing > gpurun_out/y_hit_nof.json 2> gpurun_out/y_hit_nof.err
python - <<'PY'
import json
for n in ['dsl2d','hit','hit_nof']:
    try:
        t=[l for l in open(f'gpurun_out/y_{n}.json').read().splitlines() if l.startswith('{')][-1]
        b=json.loads(t)
        print(n, round(b['ms_per_step'],2), round(b['value']/1e6,1), b['config']['mg_iters_last_step'], 'e2e', round(b['e2e']['value']/1e6,1))
    except Exception as e:
        print(n, 'failed', e); print(open(f'gpurun_out/y_{n}.err').read()[-800:])
PY
