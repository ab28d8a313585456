#!/bin/bash
# 8-GPU visit: BASELINE configs[4] HIT 512^3 on a 2x2x2 block decomposition and on z slabs (weak-scaled 256^3 per GPU)
mkdir -p gpurun_out
for D in blocks slabs; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 10 --warmup 3 --problem hit --decomp $D --no-verify \
   > gpurun_out/bench_n8_hit_$D.json 2> gpurun_out/bench_n8_hit_$D.err; grep "^{" gpurun_out/bench_n8_hit_$D.json | cut -c1-700; tail -3 gpurun_out/bench_n8_hit_$D.err
done
