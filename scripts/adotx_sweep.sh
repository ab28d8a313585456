#!/bin/bash
# tuning sweep of the marching nodal apply kernel (planes per thread, register cap)
for kb in 16 32; do for mb in 2 3; do
  echo -n "KB=$kb MINB=$mb: "; IAMRX_ADOTX_KB=$kb IAMRX_ADOTX_MINB=$mb timeout 200 python scripts/kb_time.py adotx 256 20
done; done
