#!/bin/bash
# 2-GPU: slabs and blocks with the deep-ghost cell sweeps off / on (nodal slab sweeps now exchange once per sweep)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -2
for CFG in "slabs 0" "slabs 1" "blocks 0" "blocks 1"; do
set -- $CFG
IAMRX_CELL_DEEP=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 10 --warmup 3 --decomp $1 --e2e-steps 0 \
   > gpurun_out/v_$1_$2.json 2> gpurun_out/v_$1_$2.err
python - <<PY
import json
try:
    t=[l for l in open('gpurun_out/v_$1_$2.json').read().splitlines() if l.startswith('{')][-1]
    b=json.loads(t)
    print('$1 celldeep=$2', round(b['ms_per_step'],2), round(b['value']/1e6,1), b['config']['mg_iters_last_step'], b['gpu_launches']/10, b.get('verify',{}).get('linf_state_vs_single_rank_layout'))
except Exception as e:
    print('$1 $2 failed', e); print(open('gpurun_out/v_$1_$2.err').read()[-800:])
PY
done
