#!/bin/bash
# attention exp2 split sweep (IIR_ATTN_POLY = n: every n-th exponential on the FMA pipe; 0 = all on the SFU) on the whole step
mkdir -p gpurun_out
for P in 4 0 8 3 4; do
  IIR_ATTN_POLY=$P timeout 300 python bench.py --no-cpu --no-fp16 --no-vae --steps 30 > gpurun_out/poly_$P.json 2> gpurun_out/poly.err
  python - "$P" <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/poly_{sys.argv[1]}.json').read().strip().split('\n')[-1])
kb=d['kernel_breakdown']
print('POLY', sys.argv[1], round(d['ms_per_step'],3), 'ms/step  clocks', d['clocks']['sm_mhz'], ' single-stream attn ms', round(kb['attn_tc']['ms'],3))
PY
done
