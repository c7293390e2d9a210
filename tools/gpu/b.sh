#!/bin/bash
mkdir -p gpurun_out
for b in 2 4 8; do
timeout 900 python bench.py --batch $b --steps 8 --no-cpu --no-fp16 > gpurun_out/bench_b$b.json 2> gpurun_out/bench.err; echo "b$b rc=$?"; tail -2 gpurun_out/bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_b$b.json').read().strip().split('\n')[-1])
print('B=$b', {k:d[k] for k in ('value','ms_per_step','step_tflops','step_frac_of_sustained_peak')}, d['clocks'])
PY
done
