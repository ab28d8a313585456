#!/bin/bash
# 8-GPU visit: block decomposition with the deep-ghost fused nodal sweep (TaylorGreen verify leg + HIT 512^3), slabs for reference
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n8.txt 2>&1
for CFG in "tg blocks" "hit blocks" "tg slabs"; do
set -- $CFG
V=""; [ "$1" = "hit" ] && V="--no-verify"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 10 --warmup 3 --problem $1 --decomp $2 $V \
   > gpurun_out/bench_n8_$1_$2_v2.json 2> gpurun_out/bench_n8_$1_$2_v2.err
python - <<PY
import json
try:
    t=[l for l in open('gpurun_out/bench_n8_$1_$2_v2.json').read().splitlines() if l.startswith('{')][-1]
    b=json.loads(t)
    print('$1 $2', round(b['ms_per_step'],2), round(b['value']/1e6,1), 'e2e', round(b['e2e']['value']/1e6,1), b.get('verify'), b['config']['mg_iters_last_step'], b['config'].get('host_numa_node_rank0'), 'nodal_gs', round(b['roofline']['other_kernels']['nodal_gs']['ms_per_step'],1))
except Exception as e:
    print('$1 $2 failed', e); print(open('gpurun_out/bench_n8_$1_$2_v2.err').read()[-1500:])
PY
done
