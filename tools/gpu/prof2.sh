#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --no-cpu --no-fp16 --steps 10 > gpurun_out/bench_quick.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().split('\n')[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','step_tflops')})
print('roofline',d.get('roofline'))
for k,v in sorted(d['kernel_breakdown'].items(), key=lambda kv:-kv[1]['ms']): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()})
PY
python tools/ncu_step.py > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
SKIP=$(grep -o "NSKIP=[0-9]*" gpurun_out/plain.log | cut -d= -f2)
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 300 -c 6 -o gpurun_out/prof_gemm_r01b python tools/ncu_step.py > gpurun_out/ncu2.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu2.log
