#!/bin/bash
# 1-GPU experiments: fused red+black GSRB on the HBM-resident level only; BiCGStab bottom solver; RayleighTaylor single level through bench.py
mkdir -p gpurun_out
run() { # name, env..., -- bench args
  local name=$1; shift
  env "$@" > /dev/null 2>&1
}
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0"
$B > gpurun_out/x_base.json 2> gpurun_out/x_base.err
IAMRX_GSRB_FUSED=1 IAMRX_GSRB_FUSED_TY=6 IAMRX_GSRB_FUSED_MIN=8000000 $B > gpurun_out/x_fused6.json 2> gpurun_out/x_fused6.err
IAMRX_GSRB_FUSED=1 IAMRX_GSRB_FUSED_TY=14 IAMRX_GSRB_FUSED_MIN=8000000 $B > gpurun_out/x_fused14.json 2> gpurun_out/x_fused14.err
$B --bottom-solver bicgstab > gpurun_out/x_bicg.json 2> gpurun_out/x_bicg.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --problem rt > gpurun_out/x_rt.json 2> gpurun_out/x_rt.err
python - <<'PY'
import json
for n in ['base','fused6','fused14','bicg','rt']:
    try:
        t=[l for l in open(f'gpurun_out/x_{n}.json').read().splitlines() if l.startswith('{')][-1]
        b=json.loads(t)
        print(n, round(b['ms_per_step'],2), round(b['value']/1e6,1), b['config']['mg_iters_last_step'], 'gsrb', round(b['roofline']['ms_per_step'],2), round(b['roofline']['frac'],3), 'e2e', b['e2e']['value'])
    except Exception as e:
        print(n, 'failed', e); print(open(f'gpurun_out/x_{n}.err').read()[-800:])
PY
