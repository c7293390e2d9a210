#!/bin/bash
# bench.py under torchrun on N GPUs of one box (main data-parallel line + partition sub-records), NCCL rank lines kept
N=${1:-8}
mkdir -p gpurun_out
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench N=$N rc=$?"
grep -v "NCCL INFO" gpurun_out/r02_bench_n$N.err | grep -v "^\*\|OMP_NUM" | tail -8 | cut -c1-300
python - $N <<'PY'
import json, sys
n=sys.argv[1]
d=json.loads(open(f'gpurun_out/r02_bench_n{n}.json').read().strip().split('\n')[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','step_frac_of_sustained_peak','dtype')}, d['clocks'], 'e2e', d['e2e']['value'])
for k,v in (d.get('partitions') or {}).items(): print(k, {kk:(round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('images','ms_per_step','value','e2e_value','frac_of_sustained_peak_per_gpu','skipped','error')})
PY
