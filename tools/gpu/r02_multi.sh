#!/bin/bash
# N GPUs of one box: the multi-GPU parity tests, then bench.py under torchrun (main data-parallel line + the
# communicating partitions of SURVEY §8e as sub-records)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_multi_gpu.py -q --timeout 800 -rf 2>&1 | tail -4
NCCL_DEBUG=INFO timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench N=$N rc=$?"
grep -c "NCCL INFO" gpurun_out/r02_bench_n$N.err; grep -E "NVLS|nranks|Connected all rings|comm 0x" gpurun_out/r02_bench_n$N.err | head -6; grep -v "NCCL INFO" gpurun_out/r02_bench_n$N.err | tail -5
python - $N <<'PY'
import json, sys
n=sys.argv[1]
d=json.loads(open(f'gpurun_out/r02_bench_n{n}.json').read().strip().split('\n')[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','step_frac_of_sustained_peak','dtype')}, d['clocks'], 'e2e', d['e2e']['value'])
for k,v in (d.get('partitions') or {}).items(): print(k, {kk:(round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('images','ms_per_step','value','e2e_value','frac_of_sustained_peak_per_gpu','skipped','error')})
PY
