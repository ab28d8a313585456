#!/bin/bash
# multi-GPU visit: N ranks (N = number of visible GPUs): multi-GPU parity test, weak-scaling bench with the verify leg (slabs and blocks)
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_multigpu_n$N.log; cat gpurun_out/pytest_multigpu_n$N.log
for D in slabs blocks; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --decomp $D \
   > gpurun_out/bench_n${N}_$D.json 2> gpurun_out/bench_n${N}_$D.err
python - <<PY
import json
try:
    b=json.load(open('gpurun_out/bench_n${N}_$D.json'))
    print('$D', b['n_gpus'], round(b['ms_per_step'],2), round(b['value']/1e6,1), 'e2e', round(b['e2e']['value']/1e6,1), b.get('verify'), b['config']['mg_iters_last_step'])
except Exception as e:
    print('$D failed', e); print(open('gpurun_out/bench_n${N}_$D.err').read()[-1500:])
PY
done
if [ "$N" = "8" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 10 --warmup 3 --problem hit --decomp blocks --no-verify \
   > gpurun_out/bench_n8_hit512.json 2> gpurun_out/bench_n8_hit512.err; cat gpurun_out/bench_n8_hit512.json | cut -c1-600; tail -3 gpurun_out/bench_n8_hit512.err
fi
