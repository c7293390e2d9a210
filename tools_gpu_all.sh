#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -x 2>&1 | tail -15 > gpurun_out/all_tests.log; echo "tests rc=$?"; tail -6 gpurun_out/all_tests.log
timeout 600 python tools/profile_step.py config2 > gpurun_out/shapes_config2.txt 2>&1; head -24 gpurun_out/shapes_config2.txt
timeout 1200 python bench.py --steps 30 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','step_tflops','step_frac_of_sustained_peak')})
print('e2e',d['e2e']['value'], d['e2e']['seconds_per_image_batch']); print('roofline',d['roofline']); print('clocks',d['clocks'])
PY
tail -5 gpurun_out/bench.err
