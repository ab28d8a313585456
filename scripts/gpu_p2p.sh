#!/bin/bash
# 2-GPU check of the peer-memory exchange: NCCL-vs-oracle test, then the N=2 bench lines (verify leg included)
mkdir -p gpurun_out
export IAMRX_P2P_VERBOSE=1
timeout 300 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -4
for D in slabs blocks; do
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 10 --warmup 3 --decomp $D \
   > gpurun_out/p2p_$D.json 2> gpurun_out/p2p_$D.err
echo "rc $?"; grep -h "peer-memory" gpurun_out/p2p_$D.err | head -2
python - <<PY
import json
try:
    t=[l for l in open('gpurun_out/p2p_$D.json').read().splitlines() if l.startswith('{')][-1]
    b=json.loads(t)
    print('$D', round(b['ms_per_step'],2), round(b['value']/1e6,1), 'e2e', round(b['e2e']['value']/1e6,1), b['config']['mg_iters_last_step'], b['gpu_launches']/10, b.get('verify',{}).get('linf_state_vs_single_rank_layout'))
except Exception as e:
    print('$D failed', e); print(open('gpurun_out/p2p_$D.err').read()[-1500:])
PY
done
