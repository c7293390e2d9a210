#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench.err; echo "default rc=$?"; tail -2 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().split('\n')[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','step_tflops','step_frac_of_sustained_peak','dtype')})
print('e2e',d['e2e']['value'],'fp16',d.get('fp16'),'cpu',d.get('cpu_baseline'),'roofline',d.get('roofline'),'clocks',d['clocks'])
PY
# launch list of the bench command (plain run first, same args)
python bench.py --steps 1 --warmup 3 --no-cpu --no-fp16 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-fp16 > gpurun_out/ncu.log 2>&1
echo "ncu list rc=$?"; tail -2 gpurun_out/ncu.log; wc -l gpurun_out/launches_r01b.csv
