#!/bin/bash
mkdir -p gpurun_out
if [ "$1" != "notest" ]; then
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread -x 2>&1 | tail -15 > gpurun_out/all_tests.log; echo "tests rc=$?"; tail -4 gpurun_out/all_tests.log
fi
show() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().split('\n')[-1])
print(sys.argv[1], {k:d[k] for k in ('value','ms_per_step','gpu_launches','step_tflops','step_frac_of_sustained_peak','dtype')})
print('  e2e',d['e2e']['value'], 'fp16', d.get('fp16'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'clocks', d['clocks'])
PY
}
timeout 900 python bench.py --no-cpu --no-fp16 > gpurun_out/bench_quick.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err; show gpurun_out/bench_quick.json
IIR_NO_PDL=1 timeout 600 python tools/trace_step.py config2 1 > gpurun_out/trace_config2_nopdl.txt 2>&1; echo "trace rc=$?"; grep -v Warn gpurun_out/trace_config2_nopdl.txt | head -24
