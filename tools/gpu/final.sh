#!/bin/bash
# round-end style validation: gpu tests, smoke, both bench arms
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread -x 2>&1 | tail -5
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().split('\n')[-1])
kb=d.pop('kernel_breakdown')
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','step_frac_of_sustained_peak','dtype')}, d['clocks'])
print('e2e',d['e2e']['value'],'roofline',d['roofline']['frac'], 'fp16', d.get('fp16'), 'vae', d.get('vae_decode'), 'cpu', d['cpu_baseline']['value'])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_ref.json
