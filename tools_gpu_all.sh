#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -x 2>&1 | tail -15 > gpurun_out/all_tests.log; echo "tests rc=$?"; tail -4 gpurun_out/all_tests.log
timeout 1200 python bench.py --steps 30 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().split('\n')[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','step_tflops','step_frac_of_sustained_peak')})
print('e2e',d['e2e']['value'], d['e2e']['seconds_per_image_batch']); print('roofline',d['roofline']['achieved'], d['roofline']['frac']); print('clocks',d['clocks'])
for k,v in sorted(d['kernel_breakdown'].items(), key=lambda kv:-kv[1]['ms'])[:7]: print(f"{k:16s} n={v['launches']:4d} {v['ms']:7.2f} ms  {v['tflops'] or 0:7.1f} TF/s {v['gbs'] or 0:7.1f} GB/s")
PY
tail -3 gpurun_out/bench.err
