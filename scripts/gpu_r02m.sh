#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_step.py tests/test_forcing.py tests/test_twod.py tests/test_solvers.py -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/w_tg.json 2> gpurun_out/w_tg.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --problem hit > gpurun_out/w_hit.json 2> gpurun_out/w_hit.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --problem dsl2d > gpurun_out/w_dsl2d.json 2> gpurun_out/w_dsl2d.err
python - <<'PY'
import json
for n in ['tg','hit','dsl2d']:
    try:
        t=[l for l in open(f'gpurun_out/w_{n}.json').read().splitlines() if l.startswith('{')][-1]
        b=json.loads(t)
        print(n, round(b['ms_per_step'],2), round(b['value']/1e6,1), b['config']['mg_iters_last_step'], 'e2e', round(b['e2e']['value']/1e6,1), 'launches', b['gpu_launches']/10)
    except Exception as e:
        print(n, 'failed', e); print(open(f'gpurun_out/w_{n}.err').read()[-800:])
PY
