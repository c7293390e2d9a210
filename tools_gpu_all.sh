#!/bin/bash
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().split('\n')[-1])
print(sys.argv[1], {k:d[k] for k in ('value','ms_per_step','gpu_launches','step_tflops','step_frac_of_sustained_peak','dtype')})
print('  e2e',d['e2e']['value'], 'fp16', d.get('fp16'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'clocks', d['clocks'])
PY
}
timeout 1500 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench.err; echo "default rc=$?"; tail -2 gpurun_out/bench.err; show gpurun_out/bench_default.json
timeout 900 python bench.py --workload config3 --steps 15 --no-cpu --no-fp16 > gpurun_out/bench_config3.json 2> gpurun_out/bench.err; echo "config3 rc=$?"; tail -2 gpurun_out/bench.err; show gpurun_out/bench_config3.json
timeout 900 python bench.py --batch 4 --steps 10 --no-cpu --no-fp16 > gpurun_out/bench_b4.json 2> gpurun_out/bench.err; echo "b4 rc=$?"; tail -2 gpurun_out/bench.err; show gpurun_out/bench_b4.json
