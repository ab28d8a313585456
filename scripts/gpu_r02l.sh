#!/bin/bash
# 2-GPU experiment: multigrid consolidation threshold (cells per box below which a level becomes one replicated box) on z slabs
mkdir -p gpurun_out
for T in 32768 262144 2097152; do
IAMRX_MG_CONSOLIDATE=$T timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 --no-verify --e2e-steps 0 \
   > gpurun_out/z_cons_$T.json 2> gpurun_out/z_cons_$T.err
python - <<PY
import json
try:
    t=[l for l in open('gpurun_out/z_cons_$T.json').read().splitlines() if l.startswith('{')][-1]
    b=json.loads(t)
    print('consolidate $T', round(b['ms_per_step'],2), round(b['value']/1e6,1), b['config']['mg_iters_last_step'], b['gpu_launches']/10)
except Exception as e:
    print('$T failed', e); print(open('gpurun_out/z_cons_$T.err').read()[-800:])
PY
done
