#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "folded or gemm" --timeout 300 2>&1 | tail -8
timeout 900 python -m pytest tests/test_model_parity_gpu.py -q -x --timeout 600 2>&1 | tail -8
show() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().split('\n')[-1])
print(sys.argv[1], {k:d[k] for k in ('value','ms_per_step','gpu_launches','step_tflops','step_frac_of_sustained_peak')}, d['clocks'], (d.get('roofline') or {}).get('frac'))
PY
}
timeout 600 python bench.py --no-cpu --no-fp16 > gpurun_out/bench_fold.json 2> gpurun_out/bench_fold.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_fold.err; show gpurun_out/bench_fold.json
IIR_LN_FOLD=0 timeout 600 python bench.py --no-cpu --no-fp16 > gpurun_out/bench_nofold.json 2> gpurun_out/bench_nofold.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_nofold.err; show gpurun_out/bench_nofold.json
