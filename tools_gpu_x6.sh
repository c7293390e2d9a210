#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread -x 2>&1 | tail -5
timeout 600 python bench.py --no-cpu --no-fp16 > gpurun_out/bench_pattn.json 2> gpurun_out/bench_pattn.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_pattn.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_pattn.json').read().strip().split('\n')[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','step_tflops','step_frac_of_sustained_peak')}, d['clocks'], (d.get('roofline') or {}).get('frac'), d['e2e']['value'])
kb=d['kernel_breakdown']
print('single-stream total', round(sum(v['ms'] for v in kb.values()),2))
for k,v in kb.items(): print('  ',k, v['launches'], round(v['ms'],3), v.get('tflops') and round(v['tflops']))
PY
