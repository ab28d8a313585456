#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 10 --warmup 3 --decomp slabs --e2e-steps 0 --no-verify --kernel-table \
   > gpurun_out/u_slabs_kt.json 2> gpurun_out/u_slabs_kt.err
grep -v "^\*\|OMP_NUM\|^$\|NCCL" gpurun_out/u_slabs_kt.err | head -45
