#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -2
for D in slabs blocks; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 10 --warmup 3 --decomp $D \
   > gpurun_out/t_$D.json 2> gpurun_out/t_$D.err
python - <<PY
import json
try:
    t=[l for l in open('gpurun_out/t_$D.json').read().splitlines() if l.startswith('{')][-1]
    b=json.loads(t)
    print('$D', round(b['ms_per_step'],2), round(b['value']/1e6,1), 'e2e', round(b['e2e']['value']/1e6,1), b['config']['mg_iters_last_step'], b['gpu_launches']/10, b.get('verify',{}).get('linf_state_vs_single_rank_layout'))
except Exception as e:
    print('$D failed', e); print(open('gpurun_out/t_$D.err').read()[-800:])
PY
done
